#!/bin/bash
# end-of-round record: GPU tests, smoke(), the reference arm and the full bench line on one box
set -u
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/final_tests.log 2>&1; echo "tests rc=$?"; tail -2 $O/final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/final_smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > $O/final_ref.json 2> $O/final_ref.err; echo "ref rc=$?"; tail -c 400 $O/final_ref.json
python bench.py --steps 20 --warmup 5 > $O/final_bench.json 2> $O/final_bench.err; echo "bench rc=$?"; head -c 700 $O/final_bench.json
