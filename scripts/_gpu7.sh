set -x
cd $GRAFT_REPO_ROOT
python scripts/exp_variants.py run cur nofrozen noreserve cur noreserve 2>&1 | tee gpurun_out/r2_variants_e.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 | tee gpurun_out/r2_gputests_f.log
