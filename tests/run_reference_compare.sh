#!/bin/bash
# Config 1 parity with the reference's OWN harness (tests/compare-debug-out.py, used as-is from /root/reference).
# 1. on the GPU box:  python -m pytest tests/test_gpu_decode.py -m gpu   (writes gpurun_out/<fixture>_b200.dbg from the
#    staged CUDA path through pov_ogg_vorbis_decode_memory(debug_out=...))
# 2. here (needs /root/reference + oracle/_ref/ours.bin):  bash tests/run_reference_compare.sh
# compare-debug-out.py hard-imports better_exchook (not installed, no network): a two-function stub is put on PYTHONPATH.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp -d)
mkdir -p "$TMP/stub"
printf 'def install(): pass\ndef better_exchook(*a, **k): pass\n' > "$TMP/stub/better_exchook.py"
for n in stereo44khz mono44khz; do
  "$ROOT/oracle/_ref/ours.bin" --in "/root/reference/tests/audio/test.$n.ogg" --debug_out "$TMP/$n.ref.dbg" > /dev/null
  echo "== $n: reference decoder dump vs B200 dump =="
  PYTHONPATH="$TMP/stub" python /root/reference/tests/compare-debug-out.py --ourout "$TMP/$n.ref.dbg" \
      --libvorbisout "$ROOT/gpurun_out/${n}_b200.dbg" | tail -6
done
rm -rf "$TMP"
