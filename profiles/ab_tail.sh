#!/bin/bash
# A/B of the tail-kernel changes on one box: tests + microbenchmarks + short benches with the new library, then the same
# benches with the library built from the previous commit (profiles/_ab_old_liblvae_b200.so, not tracked).
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/ab_tail_tests.log 2>&1; echo "tests rc=$?" > $O/ab_tail_summary.txt
tail -3 $O/ab_tail_tests.log >> $O/ab_tail_summary.txt
python profiles/bench_tail_kernels.py > $O/ab_tail_new.txt 2>&1; echo "tail new rc=$?" >> $O/ab_tail_summary.txt
python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --no-hbm-rooflines > $O/ab_bench_new.json 2> $O/ab_bench_new.err; echo "bench new rc=$?" >> $O/ab_tail_summary.txt
python bench.py --workload iw --steps 1 --warmup 3 --no-extras --no-cpu-baseline --no-hbm-rooflines > $O/ab_iw_new.json 2> $O/ab_iw_new.err; echo "iw new rc=$?" >> $O/ab_tail_summary.txt
if [ -f profiles/_ab_old_liblvae_b200.so ]; then
  cp ladder-vae-pytorch_b200/liblvae_b200.so /tmp/new.so
  cp profiles/_ab_old_liblvae_b200.so ladder-vae-pytorch_b200/liblvae_b200.so
  python profiles/bench_tail_kernels.py > $O/ab_tail_old.txt 2>&1; echo "tail old rc=$?" >> $O/ab_tail_summary.txt
  python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --no-hbm-rooflines > $O/ab_bench_old.json 2> $O/ab_bench_old.err; echo "bench old rc=$?" >> $O/ab_tail_summary.txt
  python bench.py --workload iw --steps 1 --warmup 3 --no-extras --no-cpu-baseline --no-hbm-rooflines > $O/ab_iw_old.json 2> $O/ab_iw_old.err; echo "iw old rc=$?" >> $O/ab_tail_summary.txt
  cp /tmp/new.so ladder-vae-pytorch_b200/liblvae_b200.so
fi
for f in new old; do
  python - <<EOF >> $O/ab_tail_summary.txt
import json
for w in ("bench", "iw"):
    try:
        d = json.loads([l for l in open("$O/ab_%s_$f.json" % w) if l.startswith("{")][-1])
        print("$f", w, "%.3f ms/step" % d["ms_per_step"], "%.1f" % d["value"], d["unit"])
    except Exception as e:
        print("$f", w, "unreadable:", e)
EOF
done
cat $O/ab_tail_summary.txt
echo "--- new"; cat $O/ab_tail_new.txt; echo "--- old"; cat $O/ab_tail_old.txt
