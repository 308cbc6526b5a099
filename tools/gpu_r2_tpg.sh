cd tests
for t in 32 16 8 4; do echo "== NGB_LU_TPG=$t"; for i in 1 2; do NGB_LU_TPG=$t timeout 120 python gpu_profile_run.py 4096 2>&1 | tail -1; done; NGB_LU_TPG=$t timeout 120 python gpu_profile_run.py 4096 stages 2>&1 | tail -1; done
