#!/bin/bash
# Times tools/probe_warp_bwd.py for every tuning build under build/.
shopt -s nullglob
for lib in default build/libtcsfm_*.so; do
  if [ "$lib" != default ]; then export TCSFM_B200_LIB=$PWD/$lib; else unset TCSFM_B200_LIB; fi
  echo "== $lib"; python tools/probe_warp_bwd.py 2>/dev/null | head -2
done
