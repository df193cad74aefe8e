timeout 300 python profiles/post_timeline.py 2>&1 | tail -30
