#!/bin/bash
# A/B builds of libpov_synth.so on the same box: tools/ab.sh [lib ...]   (default: in-tree build vs build/ab/libpov_base.so)
# prints ms per launch and the bench's own correctness check (max-abs error vs the CPU oracle, packets with status)
libs=("$@"); [ ${#libs[@]} -eq 0 ] && libs=(parseoggvorbis_b200/libpov_synth.so build/ab/libpov_base.so)
for rep in 1 2; do for l in "${libs[@]}"; do
  POV_LIB_PATH=$PWD/$l python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); c=d.get('check') or {}
    print('%.4f ms  err %.3g  snr %.1f  bad %s  $l' % (d['ms_per_step'], c.get('max_abs_err', -1), c.get('snr_db', -1), c.get('packets_with_status')))
except Exception as e: print('failed: $l', e)"
done; done
