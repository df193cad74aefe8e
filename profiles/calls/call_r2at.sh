for cfg in "B2_PBLOCK_CTAS=2 B2_PBLOCK_NT=256" "B2_PBLOCK_CTAS=1 B2_PBLOCK_NT=512 B2_PBLOCK_HOT=7" "B2_PBLOCK_CTAS=1 B2_PBLOCK_NT=1024 B2_PBLOCK_HOT=7" "B2_PBLOCK_CTAS=2 B2_PBLOCK_NT=256"; do
  echo "== $cfg (64 chains)"; env $cfg timeout 300 python profiles/sv_ncu_target.py 64 2>&1 | tail -1
done
echo "== default (148 chains)"; timeout 300 python profiles/sv_ncu_target.py 148 2>&1 | tail -1
echo "== 512 threads (148 chains)"; B2_PBLOCK_CTAS=1 B2_PBLOCK_NT=512 B2_PBLOCK_HOT=7 timeout 300 python profiles/sv_ncu_target.py 148 2>&1 | tail -1
