#!/bin/bash
# sweep the number of chunks of the host-buffer pipeline (RMB_HOST_CHUNKS) and print the e2e figure
for c in 1 2 3 4 6 8; do
  RMB_HOST_CHUNKS=$c python bench.py --no-cpu-baseline --steps 5 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('chunks', $c, round(d['e2e']['value']))"
done
