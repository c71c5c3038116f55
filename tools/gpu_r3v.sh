#!/bin/bash
for g in 1 2 4; do python tools/multi_handle.py 12 4096 $g 6 0 2>&1 | tail -1; done
python tools/multi_handle.py 12 4096 2 6 108 2>&1 | tail -1
python tools/multi_handle.py 12 4096 4 6 54 2>&1 | tail -1
python tools/multi_handle.py 12 8192 2 4 108 2>&1 | tail -1
python tools/multi_handle.py 12 8192 1 4 0 2>&1 | tail -1
