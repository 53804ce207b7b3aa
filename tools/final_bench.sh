#!/bin/bash
# Bench lines + ncu captures of the batched (tcgen05) path, as committed under profiles/ (run on a B200 through gpurun).
set -u
O=gpurun_out
python bench.py --batch 256 --frames 64 --prompt-bytes 64 --steps 3 --warmup 3 --no-cpu-baseline > $O/g_bs256.json 2> $O/g_bench.err
python bench.py --batch 256 --frames 128 --steps 2 --warmup 2 --no-cpu-baseline > $O/g_bs256_p200.json 2>> $O/g_bench.err
python bench.py --batch 64 --frames 128 --sampled --steps 3 --warmup 3 --no-cpu-baseline > $O/g_bs64.json 2>> $O/g_bench.err
timeout 600 python bench.py --batch 32 --frames 4096 --steps 1 --warmup 1 --no-cpu-baseline > $O/g_bs32_long.json 2>> $O/g_bench.err
CMD="python bench.py --batch 256 --frames 8 --prompt-bytes 64 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > $O/plain_ncu_cmd3.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/g_launches_bs256.csv $CMD > $O/ncu_list3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:smol_decode_kernel -s 3 -c 1 -o $O/tc_full_final2 $CMD > $O/ncu_full_tc3.log 2>&1
timeout 200 python tools/phase_profile.py --batch 256 --frames 16 --prompt-bytes 64 > $O/phase_tc_bs256_final.log 2>&1
timeout 200 python tools/phase_profile.py --batch 64 --frames 16 --sampled > $O/phase_tc_bs64_final.log 2>&1
python - <<PY
import json
for f in ["g_bs256","g_bs256_p200","g_bs64","g_bs32_long"]:
    try:
        d=json.loads(open("gpurun_out/"+f+".json").read().strip().splitlines()[-1])
        print(f, round(d["value"],1), "frames/s", round(d["config"]["us_per_frame"]), "us/step", d.get("roofline",{}).get("bound"), round(d.get("roofline",{}).get("frac",0),4), "e2e", round(d["e2e"]["value"],1), d.get("clocks",{}).get("reasons"))
    except Exception as e: print(f, "ERR", e)
PY
tail -2 $O/ncu_full_tc3.log
