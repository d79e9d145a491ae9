#!/bin/bash
for BW in 4 8 16; do SKM_K2_BW=$BW timeout 200 python tools/probe_k2.py 100 784 78 5000000 2>&1 | tail -1; done
for BW in 8 16 32; do SKM_K2_BW=$BW timeout 200 python tools/probe_k2.py 32 1024 51 8000000 2>&1 | tail -1; done
for BW in 4 8; do SKM_K2_BW=$BW timeout 200 python tools/probe_k2.py 128 512 40 5000000 2>&1 | tail -1; done
