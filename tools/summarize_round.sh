#!/bin/bash
# Turns what tools/collect_profiles.sh left in gpurun_out/ into the committed files under profiles/ (run here, after the
# GPU call has come back).   usage: tools/summarize_round.sh r02
set -e
R=${1:-r02}
O=gpurun_out
P=profiles
cp $O/${R}_bench.json $P/${R}_bench.json
cp $O/${R}_bench_reference_arm.json $P/${R}_bench_reference_arm.json
cp $O/${R}_launches.csv $P/${R}_ncu_launch_list.csv
python $P/summarize_ncu.py $O/${R}_prof_warp.ncu-rep > $P/${R}_ncu_k_warp.txt
python $P/summarize_ncu.py $O/${R}_prof_decode.ncu-rep > $P/${R}_ncu_k_decode.txt
python $P/summarize_ncu.py $O/${R}_prof_residue.ncu-rep > $P/${R}_ncu_k_residue.txt
python $P/stalls_by_function.py $O/${R}_prof_warp.ncu-rep parseoggvorbis_b200/csrc/kernel_warp.cu > $P/${R}_ncu_k_warp_synth_stalls_by_function.txt
[ -s $O/${R}_configs.jsonl ] && cp $O/${R}_configs.jsonl $P/${R}_configs_3_4_kernel_times.jsonl
[ -s $O/${R}_features.json ] && cp $O/${R}_features.json $P/${R}_config4_feature_throughput.json
[ -s $O/${R}_pcie_probe.log ] && cp $O/${R}_pcie_probe.log $P/${R}_pcie_probe.log
[ -s $O/${R}_tmem_probe_throughput.log ] && cp $O/${R}_tmem_probe_throughput.log $P/${R}_tmem_probe_throughput.log
# DRAM traffic of the captured k_warp_synth launch next to the algorithmic bytes (bench.py reports it as roofline.traffic)
python - "$P/${R}_ncu_k_warp.txt" "$P/${R}_bench.json" "$P/${R}_traffic.json" "$R" <<'PY'
import json, re, sys
txt, bench, out, rnd = sys.argv[1:5]
val = {}
for line in open(txt):
    m = re.match(r"\s*(dram__bytes_(?:read|write)\.sum)\s+([0-9.]+)\s+(\w+)", line)
    if m:
        val[m.group(1)] = float(m.group(2)) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[m.group(3)]
b = json.loads(open(bench).read().strip().splitlines()[-1])
alg = b["roofline"]["algorithmic_bytes_per_launch"]
rd, wr = val["dram__bytes_read.sum"], val["dram__bytes_write.sum"]
json.dump({"kernel": "k_warp_synth", "samples_per_launch": b["samples_per_step_per_gpu"], "traffic_bytes_per_launch": rd + wr,
           "dram_read_bytes": rd, "dram_write_bytes": wr, "algorithmic_bytes_per_launch": alg, "ratio": (rd + wr) / alg,
           "source": "ncu --set full, gpurun_out/%s_prof_warp.ncu-rep (profiles/%s_ncu_k_warp.txt), bench.py --steps 3 --warmup 3 "
                     "--no-cpu-baseline --no-e2e" % (rnd, rnd)}, open(out, "w"), indent=1)
PY
echo "profiles/${R}_* refreshed"
