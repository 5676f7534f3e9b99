set -x
mkdir -p gpurun_out
( time python -m pytest tests -x -q -m gpu ) > gpurun_out/final_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/final_smoke.log
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?" >> gpurun_out/final_bench.err
python bench.py --impl reference --steps 2 > gpurun_out/final_bench_ref.json 2>> gpurun_out/final_bench.err
python bench.py --no-cpu --no-e2e --no-sweep --front-end-sweep --steps 5 > gpurun_out/final_sweep.json 2>> gpurun_out/final_bench.err
tail -3 gpurun_out/final_tests.log; cat gpurun_out/final_smoke.log | tail -2; cut -c1-400 gpurun_out/final_bench.json
