cd /root/repo
for B in 8 6 5 4; do
  MV_NVCC_EXTRA=-DMV_TILE_BLOCKS=$B python -c "
from megaverse_b200 import _build; _build.build_lib(force=True, verbose=True)" 2>&1 | grep -A2 "tileKernelILb1" | grep -E "spill|registers" | tr '\n' ' '
  echo " <- blocks/SM $B"
  python tools/submit_rate.py 2>&1 | tail -2
done
