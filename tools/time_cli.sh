#!/bin/bash
# tools/time_cli.sh -- wall time of the reference's config 1 (`ekgSim test -sim <16 params> -out result`) through the B200 CLI,
# process start to exit, without and with the binary shape cache.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
D=$(mktemp -d)
python -c "import sys; sys.path.insert(0, '$ROOT/tests'); import ekgio; ekgio.materialise_testrun('$D')"
cd "$D"
P=0.00035813,0.0890636,0.0632915,226.183,0.000369406,0.0965625,0.0523254,232.278,0.000710767,0.0720323,0.0187579,200.93,23,22,15,13
TIMEFORMAT="process wall %R s"
for i in 1 2; do
  { time "$ROOT/ekgsim_b200/bin/ekgSim" test -sim $P -out result > out.txt 2> err.txt; } 2>&1
  grep -h "simulation done\|criteria" out.txt
done
grep -h "done (" err.txt | tr '\n' ';'; echo
export EKGSIM_B200_CACHE=1
{ time "$ROOT/ekgsim_b200/bin/ekgSim" test -sim $P > out.txt 2> err.txt; } 2>&1
{ time "$ROOT/ekgsim_b200/bin/ekgSim" test -sim $P > out.txt 2> err.txt; } 2>&1
grep -h "loading shape" err.txt
