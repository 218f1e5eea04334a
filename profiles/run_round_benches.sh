#!/bin/bash
# All bench lines of a round on one B200 (gpurun): bash profiles/run_round_benches.sh TAG
TAG=${1:-vX}
O=gpurun_out
python bench.py > $O/bench_${TAG}.json 2> $O/bench_${TAG}.err
python bench.py --impl reference --steps 6 --warmup 1 > $O/bench_${TAG}_ref.json 2>> $O/bench_${TAG}.err
for wl in hss_td3 mass_td1 mass_td3_mixed unsafe_td1; do
  python bench.py --workload $wl --skip-cpu > $O/bench_${TAG}_${wl}.json 2>> $O/bench_${TAG}.err
done
python bench.py --workload mass_td3_srew --policy --skip-cpu > $O/bench_${TAG}_policy.json 2>> $O/bench_${TAG}.err
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench_${TAG}*.json")):
    try:
        d = json.load(open(f))
        print("%-44s value %.3e  ms/step %8.3f  e2e %.3e  roofline %.4f" % (f.split("/")[-1], d["value"], d.get("ms_per_step", 0), d["e2e"]["value"], d.get("roofline", {}).get("frac", 0)))
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
tail -3 $O/bench_${TAG}.err
