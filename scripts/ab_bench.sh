# usage: ab_bench.sh "lib1.so lib2.so ..."  ("default" = the in-tree library): bench.py kernel value per library
for v in $1; do
  if [ "$v" = default ]; then unset CITYSEER_B200_LIB; else export CITYSEER_B200_LIB=$PWD/$v; fi
  echo -n "== $v: "; python bench.py --no-cpu --steps 5 --warmup 3 2>/dev/null | grep -o "\"value\": [0-9.]*\|\"kernel_ms_per_step\": [0-9.]*" | tr "\n" " "; echo
done
