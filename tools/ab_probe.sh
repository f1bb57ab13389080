#!/bin/bash
# A/B timing of library builds on one GPU box: tools/ab_probe.sh <workload> <lib.so> [<lib.so> ...]
# (each lib is loaded through RMB_LIB; "default" = the in-tree build)
wl=$1; shift
for lib in "$@"; do
  if [ "$lib" = default ]; then unset RMB_LIB; else export RMB_LIB=$PWD/$lib; fi
  echo "== $lib ($wl)"
  python tools/matvec_probe.py $wl | tail -2
  python bench.py --workload $wl --steps ${STEPS:-60} --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'mv us', round(d['roofline']['avg_launch_us'],1), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value']), 'mv share', round(d['roofline']['share_of_step'],3), 'sm MHz', d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
