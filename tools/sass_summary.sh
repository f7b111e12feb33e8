#!/bin/bash
# SASS evidence per object (no GPU needed): counts of the tcgen05 / TMA / mma.sync mnemonics in the
# sm_100a code of every compiled source.  usage: tools/sass_summary.sh > profiles/r2_sass_summary.txt
cd "$(dirname "$0")/../audio_diffusion_pytorch_b200/build" || exit 1
echo "# cuobjdump -sass <object> | grep -c <mnemonic>   (nvcc $(nvcc --version | grep release | sed 's/.*release //'), -gencode arch=compute_100a,code=sm_100a)"
printf "%-18s %9s %9s %9s %9s %9s %9s %9s %9s %9s\n" object UTCHMMA UTCHMMA.2CTA LDTM STTM UTMALDG UTMALDG.2CTA UTCBAR HMMA LDSM
for o in *.o; do
  s=$(cuobjdump -sass "$o" 2>/dev/null)
  c() { echo "$s" | grep -c "$1"; }
  printf "%-18s %9s %9s %9s %9s %9s %9s %9s %9s %9s\n" "$o" "$(c 'UTCHMMA')" "$(c 'UTCHMMA.2CTA')" "$(c 'LDTM')" "$(c 'STTM')" \
    "$(c 'UTMALDG')" "$(c 'UTMALDG.*2CTA')" "$(c 'UTCBAR')" "$(c ' HMMA')" "$(c 'LDSM')"
done
echo
echo "# kernels (entry points) per object and their register / shared-memory use (ptxas -v)"
for l in *.cu.log; do
  echo "## ${l%.log}"
  grep -E "Compiling entry function|Used [0-9]+ registers" "$l" | sed 's/ptxas info    : //' | paste - - | \
    sed -E "s/Compiling entry function '([^']*)' for 'sm_100a'/\1/" | awk '{print "  " $0}' | cut -c1-220 | head -60
done
