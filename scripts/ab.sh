#!/bin/bash
# A/B builds of the CUDA library on the same box: usage scripts/ab.sh lib1.so lib2.so ... (paths relative to repo root)
# env AB_SCENES overrides the scene list; env AB_KNOBS the knob dict passed to sweep_knobs.py
for scene in ${AB_SCENES:-hairball kitchen}; do
  for lib in "$@"; do
    echo -n "$lib  "; TRAY_CUDA_LIB=$PWD/$lib python scripts/sweep_knobs.py $scene "${AB_KNOBS:-{\"TRAY_CUDA_TRI_WEIGHT\": [4], \"TRAY_CUDA_REFILL_MIN\": [4]\}}" 2>&1 | tail -1
  done
done
