echo "== current, EPI1, 6 stages"; B2_TC_EPI=1 timeout 300 python profiles/lockstep_profile.py 2>&1 | tail -1
echo "== current, EPI1, 5 stages"; B2_TC_EPI=1 B2_TC_STAGES_UNFUSED=5 timeout 300 python profiles/lockstep_profile.py 2>&1 | tail -1
echo "== current, EPI0, 5 stages"; B2_TC_EPI=0 B2_TC_STAGES_UNFUSED=5 timeout 300 python profiles/lockstep_profile.py 2>&1 | tail -1
