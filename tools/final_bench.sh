#!/bin/bash
# Round-2 measurement set, as committed under profiles/ (run on a B200 box through gpurun; ~10 GPU-minutes):
#   bench lines (ours + the CPU reference arm), ncu launch list of the bench command, ncu --set full of the data-flow
#   kernel (16 frames) and of the tcgen05 kernel (bs=256, 8 frames), DRAM traffic of the bench's own 1024-frame launch and of
#   a bs=256 launch, phase tables, the data-flow kernel's cycle / skew trace and the team timings.
set -u
O=gpurun_out
T=${1:-r2z}
python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err
python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err
# launch list of the same command (short form): per-launch times are cold-cache and serialised, the SHARE must agree
CMD="python bench.py --steps 2 --warmup 3 --configs none --no-cpu-baseline"
$CMD > $O/${T}_plain_list.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_bench.csv $CMD > $O/${T}_ncu_list.log 2>&1
# data-flow kernel: full set on a 16-frame launch; DRAM bytes of one 1024-frame launch at the bench's context
python tools/ll_ncu.py --frames 16 > $O/${T}_ll_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:smol_ll2_kernel -s 1 -c 1 -o $O/${T}_ll2 python tools/ll_ncu.py --frames 16 > $O/${T}_ll_ncu.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:smol_ll2_kernel -s 1 -c 1 --csv --log-file $O/${T}_ll2_traffic_1024.csv python tools/ll_ncu.py --frames 1024 > $O/${T}_ll_traffic.log 2>&1
# tcgen05 kernel at bs=256: full set on an 8-frame launch (the decode launch after prefill + warm-up), DRAM bytes
TC="python bench.py --batch 256 --frames 8 --prompt-bytes 64 --steps 1 --warmup 3 --configs none --no-cpu-baseline"
$TC > $O/${T}_tc_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:smol_decode_kernel -s 6 -c 2 -o $O/${T}_tc $TC > $O/${T}_tc_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:smol_decode_kernel -c 12 --csv --log-file $O/${T}_launches_bs256.csv $TC > $O/${T}_tc_list.log 2>&1
timeout 300 python tools/phase_profile.py --batch 256 --frames 16 --prompt-bytes 64 > $O/${T}_phase_bs256.log 2>&1
timeout 300 python tools/phase_profile.py --batch 256 --frames 16 --prompt-bytes 400 > $O/${T}_phase_bs256_L420.log 2>&1
timeout 300 python tools/phase_profile.py --batch 64 --frames 16 --sampled > $O/${T}_phase_bs64_sampled.log 2>&1
timeout 400 python tools/ll2_probe.py --skip-check --holdoffs 400 --staggers 0 > $O/${T}_probe.log 2>&1
timeout 400 python tools/ll2_probe.py --skip-check --skip-timing --teams > $O/${T}_teams.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1
tail -3 $O/${T}_bench.json | cut -c1-600
