#!/bin/bash
# Dev aid: one --set full capture of the 4 KiB-page launch (where does a small page spend its time)
O=gpurun_out/${1:-r2s}
SWEEP_SINGLE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:bgx_decode_pages -c 1 -f -o ${O}_page4k python scripts/page_size_sweep.py 1024 mixed 4096 > ${O}_ncu_page4k.log 2>&1
tail -2 ${O}_ncu_page4k.log
