set -x
cd $GRAFT_REPO_ROOT
python scripts/exp_variants.py run cur ctas6 ctas4 cur ctas6 2>&1 | grep -v "^+" | tee gpurun_out/r2_variants_ctas.log
