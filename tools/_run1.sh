set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 200 python tools/lane_ab.py bpe32k:en --variants "CR=1;CR=0;CR=1,C=0" --reps 5 2>&1 | tail -3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:encode_bpe_lane2 --launch-skip 2 -c 1 -f -o gpurun_out/r02b_bpe_lane2 python tools/lane_ab.py bpe32k:en --variants "CR=1" --reps 3 > gpurun_out/ncu_bpe.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:encode_unigram_lane_kernel --launch-skip 2 -c 1 -f -o gpurun_out/r02b_unigram_lane python tools/lane_ab.py uni32k:en --variants "FW=1" --reps 3 > gpurun_out/ncu_uni.log 2>&1
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"encode_unigram_lane_plain" --launch-skip 1 -c 1 --csv --log-file gpurun_out/traffic_mixed.csv python tools/lane_ab.py mix_bf8k:mixed --variants "D=1" --reps 2 > /dev/null 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_all_n1.json 2> gpurun_out/r02_bench_all_n1.err; tail -3 gpurun_out/r02_bench_all_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm_n1.json 2> gpurun_out/r02_ref.err; tail -2 gpurun_out/r02_ref.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_all_n1.json") if l.startswith("{")][-1])
def show(name,r):
    print(name, "value %.1fM"%(r["value"]/1e6), ("e2e %.1fM"%(r["e2e"]["value"]/1e6)) if r.get("e2e") else "", "ms/step", round(r["ms_per_step"],3), "parity", (r.get("parity") or {}).get("result"), "roofline", (r.get("roofline") or {}).get("frac"))
show("headline", d)
for k,v in d.get("workloads",{}).items(): show(k,v)
print((d["workloads"]["bpe32k_en"].get("e2e_variants") or {}).get("warm_word_cache"))
PY
