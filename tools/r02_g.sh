#!/bin/bash
for n in 4 12 22 64; do timeout 600 python tools/matvec_probe.py h2s $n 3 2>&1 | tail -1; done
timeout 900 python tools/batch_sweep.py h2s 64 22 12 4 2>&1 | tail -8
timeout 600 python tools/batch_sweep.py h2o 500 168 84 2>&1 | tail -6
