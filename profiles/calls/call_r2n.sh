echo "== EPI0, 5 stages"; B2_TC_EPI=0 B2_TC_STAGES_UNFUSED=5 timeout 300 python profiles/lockstep_profile.py 2>&1 | tail -1
echo "== EPI0, 5 stages, smem padded by 33 KB (228 KB carve-out)"; B2_TC_EPI=0 B2_TC_STAGES_UNFUSED=5 B2_TC_SMEM_PAD=33000 timeout 300 python profiles/lockstep_profile.py 2>&1 | tail -1
echo "== EPI0, 5 stages, smem padded by 20 KB (still 196 KB carve-out)"; B2_TC_EPI=0 B2_TC_STAGES_UNFUSED=5 B2_TC_SMEM_PAD=20000 timeout 300 python profiles/lockstep_profile.py 2>&1 | tail -1
