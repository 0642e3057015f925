# usage: ab_variants.sh "lib1.so lib2.so ..." [probe args]  ("default" = the in-tree library)
# default probe: cfg4, 65 536 random sources, centrality_shortest
libs=$1; shift
args=${@:---cfg cfg4 --nsrc 65536 --reps 3}
for v in $libs; do
  echo "== variant: $v"
  if [ "$v" = default ]; then unset CITYSEER_B200_LIB; else export CITYSEER_B200_LIB=$PWD/$v; fi
  python scripts/probe.py $args 2>&1 | grep -o '"rep".*"src_per_s_kernel": [0-9.]*\|Error.*'
done
