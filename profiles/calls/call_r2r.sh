echo "== 3 stages"; B2_TC_STAGES_UNFUSED=3 timeout 300 python profiles/lockstep_profile.py 2>&1 | tail -2
echo "== 4 stages (default)"; timeout 300 python profiles/lockstep_profile.py 2>&1 | tail -2
