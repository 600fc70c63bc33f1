#!/bin/bash
for eo in 1 0; do for shp in "4096 14336" "512 14336"; do echo -n "SM_ROW_EO=$eo $shp: "; SM_ROW_EO=$eo python tools/profile_one.py $shp 6 | tail -1; done; done
SM_ROW_EO=1 timeout 600 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider -k "full_size or mid_size" 2>&1 | tail -2
