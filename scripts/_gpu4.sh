set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 | tee gpurun_out/r2_gputests_d.log
python scripts/exp_inv_variants.py run 2>&1 | tee gpurun_out/r2_inv_variants_b.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; tail -3 gpurun_out/r2_bench_d.err
timeout 300 python bench.py --config C4 --steps 3 --warmup 2 > gpurun_out/r2_bench_c4.json 2>> gpurun_out/r2_bench_d.err
timeout 300 python bench.py --config C5 --steps 3 --warmup 2 > gpurun_out/r2_bench_c5.json 2>> gpurun_out/r2_bench_d.err
timeout 300 python bench.py --config C3 --problems 1000000 --steps 3 --warmup 2 > gpurun_out/r2_bench_c3.json 2>> gpurun_out/r2_bench_d.err
