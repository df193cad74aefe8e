for cfg in "0 16" "1 16" "1 12" "1 20" "1 8" "0 16" "1 16"; do
  set -- $cfg
  echo "== B2_TC_PINGPONG=$1 B2_TC_POST_SMS=$2"
  B2_TC_PINGPONG=$1 B2_TC_POST_SMS=$2 timeout 300 python profiles/lockstep_profile.py 2>&1 | tail -2
done
B2_TC_PINGPONG=1 timeout 900 python -m pytest tests/test_gpu_parity_full_size.py tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -q --timeout 600 -x 2>&1 | tail -5
