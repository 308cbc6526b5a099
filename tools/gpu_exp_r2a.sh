cd tests
for v in cta128 cta64; do echo "== $v"; NGB200_LIB=$PWD/../build/variants/$v/libngb200.so timeout 120 python gpu_profile_run.py 4096 | tail -1; done
echo "== in-tree (asm unroll)"; timeout 120 python gpu_profile_run.py 4096 stages | tail -2
echo "== dv0 in-tree"; NGB_DV0=1 timeout 120 python gpu_profile_run.py 4096 stages | tail -2
echo "== dv0 instshared"; NGB_DV0=1 NGB200_LIB=$PWD/../build/variants/instshared/libngb200.so timeout 120 python gpu_profile_run.py 4096 stages | tail -2
