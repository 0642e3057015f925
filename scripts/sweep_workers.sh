# usage: sweep_workers.sh CFG NSRC DISTANCES "W1 W2 ..."   (prints sources/s per resident-warp count)
for w in $4; do
  echo "== workers $w"
  python scripts/probe.py --cfg $1 --nsrc $2 --reps 2 --distances $3 --workers $w 2>&1 | grep -o '"rep".*"gteps": [0-9.]*\|Error.*' | cut -c1-400
done
