# usage: ab_variants.sh "lib1.so lib2.so ..."  ("default" = the in-tree library): cfg4 probe, 65 536 random sources
for v in $1; do
  echo "== variant: $v"
  if [ "$v" = default ]; then unset CITYSEER_B200_LIB; else export CITYSEER_B200_LIB=$PWD/$v; fi
  python scripts/probe.py --cfg cfg4 --nsrc 65536 --reps 3 2>&1 | grep -o '"rep".*"src_per_s_kernel": [0-9.]*\|Error.*'
done
