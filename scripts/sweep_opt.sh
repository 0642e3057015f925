# usage: sweep_opt.sh LIB OPTION "v1 v2 ..." [probe args]  — cfg4 probe per value of a device option
lib=$1; opt=$2; vals=$3; shift 3
args=${@:---cfg cfg4 --nsrc 65536 --reps 2}
if [ "$lib" != default ]; then export CITYSEER_B200_LIB=$PWD/$lib; fi
for v in $vals; do
  echo "== $opt=$v"
  python scripts/probe.py $args --opt $opt=$v 2>&1 | grep '"rep": 1' | grep -o '"src_per_s_kernel": [0-9.]*\|"phase_kcyc": \[[^]]*\]\|"p1_[a-z]*": [0-9.]*\|"relax_per_settled": [0-9.]*' | tr '\n' ' '; echo
done
