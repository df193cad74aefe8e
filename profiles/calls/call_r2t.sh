timeout 600 python profiles/tcw_ab.py 2>&1 | tail -6
