echo "== default (64 chains)"; timeout 300 python profiles/sv_ncu_target.py 64 2>&1 | tail -1
echo "== default (148 chains)"; timeout 300 python profiles/sv_ncu_target.py 148 2>&1 | tail -1
echo "== default (296 chains)"; timeout 300 python profiles/sv_ncu_target.py 296 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "vol or sv or persistent or block or c4 or parity" 2>&1 | tail -3
