#!/bin/bash
# Round evidence in one GPU call (run through gpurun from the repo root): bench lines, ncu launch list, ncu --set full
# captures of the three kernels that matter, the probes. Everything lands in gpurun_out/; tools/summarize_round.sh turns
# it into the committed files under profiles/.   usage: tools/collect_profiles.sh r02
set -x
R=${1:-r02}
O=gpurun_out
mkdir -p $O
python bench.py > $O/${R}_bench.json 2> $O/${R}_bench.err
python bench.py --impl reference --steps 3 > $O/${R}_bench_reference_arm.json 2> $O/${R}_bench_reference_arm.err
# launch list of a short run of the same command (no numbers are taken from runs under the profiler)
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/${R}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e-dense --corpus-files 1024 --corpus-steps 1 > $O/${R}_ncu_ll.log 2>&1
# full captures: the synthesis kernel on the bench workload; the entropy-decode and residue kernels on a corpus chunk
ncu --set full --clock-control none --import-source on -k regex:k_warp -s 3 -c 1 -f -o $O/${R}_prof_warp \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/${R}_ncu_warp.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_packet_decode -s 2 -c 1 -f -o $O/${R}_prof_decode \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e-dense --corpus-files 1024 --corpus-steps 1 > $O/${R}_ncu_decode.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_residue_apply -s 2 -c 1 -f -o $O/${R}_prof_residue \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e-dense --corpus-files 1024 --corpus-steps 1 > $O/${R}_ncu_residue.log 2>&1
python tools/measure_configs.py > $O/${R}_configs.jsonl 2> $O/${R}_configs.err
python tools/measure_features.py > $O/${R}_features.json 2> $O/${R}_features.err
python tools/probes/pcie_probe.py > $O/${R}_pcie_probe.log 2>&1
tools/probes/tmem_probe 2>&1 | grep -E "^trip|^bw" > $O/${R}_tmem_probe_throughput.log
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem --format=csv > $O/${R}_gpu.txt
