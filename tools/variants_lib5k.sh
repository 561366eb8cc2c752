#!/bin/bash
for lib in build/variants/*.so; do echo "== $lib"; PISB_LIB=$PWD/$lib timeout 300 python tools/variants_one.py 3 2 1 100 5 40 2>&1 | grep -oE '"ms_per_step": [0-9.]+|"integrate": [0-9.]+' | tr '\n' ' '; echo; done
