#!/bin/bash
echo "two buffers (W<=8)"; timeout 300 python tools/probe_fwht2.py 2048 4096 8192
echo "one buffer"; SKM_FWHT_TMA_NBUF=1 timeout 300 python tools/probe_fwht2.py 2048 4096 8192
