#!/bin/bash
# Dev: A/B of runtime knobs on the bench scenes; run on a GPU box.  usage: scripts/narrow_sweep.sh '{"TRAY_CUDA_LOOKAHEAD": [0, 1, 2, 3, 5, 7, 0]}'
for scene in ${AB_SCENES:-hairball kitchen sanmiguel}; do
  python scripts/sweep_knobs.py $scene "${1:-{\"TRAY_CUDA_NARROW_MAX\": [0, 4, 0, 4]\}}" 2>&1 | grep primary
done
