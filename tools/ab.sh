#!/bin/bash
# A/B two builds of libpov_synth.so on the same box: tools/ab.sh [lib ...]   (default: in-tree build vs build/ab/libpov_base.so)
libs=("$@"); [ ${#libs[@]} -eq 0 ] && libs=(parseoggvorbis_b200/libpov_synth.so build/ab/libpov_base.so)
for rep in 1 2; do for l in "${libs[@]}"; do
  POV_LIB_PATH=$PWD/$l python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], '$l')"
done; done
