for pol in 0 1 2 0 1; do
  echo "== B2_TC_XPOLICY=$pol"; B2_TC_XPOLICY=$pol timeout 300 python profiles/lockstep_profile.py 2>&1 | tail -2
done
