#!/bin/bash
# one GPU call: the GPU test suite, the tail-kernel microbenchmarks and short train / IW benches (summary in gpurun_out/<tag>_summary.txt)
set -u
T=${1:-ab}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/${T}_tests.log 2>&1; echo "tests rc=$?" > $O/${T}_summary.txt
tail -3 $O/${T}_tests.log >> $O/${T}_summary.txt
python profiles/bench_tail_kernels.py > $O/${T}_tail.txt 2>&1; echo "tail rc=$?" >> $O/${T}_summary.txt
cat $O/${T}_tail.txt >> $O/${T}_summary.txt
python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --no-hbm-rooflines > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?" >> $O/${T}_summary.txt
python bench.py --workload iw --steps 1 --warmup 3 --no-extras --no-cpu-baseline --no-hbm-rooflines > $O/${T}_iw.json 2> $O/${T}_iw.err; echo "iw rc=$?" >> $O/${T}_summary.txt
python - <<EOF >> $O/${T}_summary.txt
import json
for w in ("bench", "iw"):
    try:
        d = json.loads([l for l in open("$O/${T}_%s.json" % w) if l.startswith("{")][-1])
        print(w, "%.3f ms/step" % d["ms_per_step"], "%.1f" % d["value"], d["unit"], "launches/step", d.get("launches_per_step"))
    except Exception as e:
        print(w, "unreadable:", e)
EOF
cat $O/${T}_summary.txt
