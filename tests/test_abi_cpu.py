"""CPU-side checks of the drop-in boundary: the C-ABI library builds/loads and exports every
symbol that include/adp_b200.h declares (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "adp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(adp_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_hot_path_entry_points():
    syms = declared_symbols()
    for name in ("adp_conv_gemm", "adp_gn_silu", "adp_ln_film", "adp_attention",
                 "adp_stem_in", "adp_stem_out", "adp_narrow_conv", "adp_sampler_step",
                 "adp_skinny_linear", "adp_time_features"):
        assert name in syms


def test_library_exports_every_declared_symbol():
    from audio_diffusion_pytorch_b200 import _build
    path = _build.build()
    lib = ctypes.CDLL(path)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in include/adp_b200.h but not exported: {missing}"
    lib.adp_version.restype = ctypes.c_int
    assert lib.adp_version() >= 1


def test_ctypes_structs_match_header_field_order():
    from audio_diffusion_pytorch_b200 import _lib
    text = open(os.path.join(ROOT, "include", "adp_b200.h")).read()
    for struct, cls in (("adp_conv_gemm_args", _lib.ConvGemmArgs),
                        ("adp_stem_in_args", _lib.StemInArgs),
                        ("adp_stem_out_args", _lib.StemOutArgs),
                        ("adp_narrow_conv_args", _lib.NarrowConvArgs)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), text, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                names.append(re.sub(r"\[.*\]", "", part.strip().split()[-1].lstrip("*")))
        assert names == [f[0] for f in cls._fields_], (struct, names)
