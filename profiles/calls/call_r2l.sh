mkdir -p gpurun_out
(cd _r1_snapshot && python -m pymc3_b200.build > /dev/null 2>&1 && echo "== r1 code" && timeout 300 python profiles/lockstep_profile.py 2>&1 | tail -3)
echo "== current, EPI0"; B2_TC_EPI=0 timeout 300 python profiles/lockstep_profile.py 2>&1 | tail -3
echo "== current, EPI1"; B2_TC_EPI=1 timeout 300 python profiles/lockstep_profile.py 2>&1 | tail -3
echo "== current, EPI1 no graph"; B2_GRAPH=0 timeout 300 python profiles/lockstep_profile.py 2>&1 | tail -3
