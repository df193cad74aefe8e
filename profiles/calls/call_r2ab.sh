for v in BASE OLDADDR NOCOPY2 PLAST SOLO BASE; do
  cp _variants/lib_$v.so pymc3_b200/libb200nuts.so
  echo "== $v"; timeout 300 python profiles/lockstep_profile.py 2>&1 | tail -1
done
