#!/bin/bash
# tests/golden/run_reference.sh -- runs the UNMODIFIED reference (oracle/_ref, built from
# /root/reference by `make -C oracle ref`) on the testRun inputs to produce the raw dumps that
# tests/golden/make_fixtures.py turns into the committed golden fixtures.  Build container only.
# About 4 core-minutes per full-length vector; everything runs in parallel in .scratch/golden/.
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
REF=/root/reference/testRun
R=$ROOT/oracle/_ref/ref_dump
G=$ROOT/.scratch/golden
mkdir -p "$G" && cd "$G"
for d in act full1 full2 full3 glue len16 v6full; do
	mkdir -p $d
	for f in simulator.ini model_24.matrix conduction_24.matrix model_24_measuring_pos.txt target_ecg_v2_v5.column target_ecg_v2_v6.column; do
		ln -sf $REF/$f $d/$f
	done
done
# second golden: v6 target
rm v6full/simulator.ini
sed 's/targets filename = target_ecg_v2_v5.column/targets filename = target_ecg_v2_v6.column/' $REF/simulator.ini > v6full/simulator.ini
echo "0.00035813,0.0890636,0.0632915,226.183,0.000369406,0.0965625,0.0523254,232.278,0.000710767,0.0720323,0.0187579,200.93,23,22,15,13" > full1/vec.txt  # README.md:108
echo "0.0005,0.05,0.05,320,0.0006,0.06,0.04,340,0.0007,0.07,0.03,310,-10,5,20,-30" > v6full/vec.txt
sed -n 1p "$ROOT/tests/golden/vectors256.txt" > full2/vec.txt
sed -n 2p "$ROOT/tests/golden/vectors256.txt" > full3/vec.txt
cp "$ROOT/tests/golden/vectors256.txt" glue/vec.txt
sed -n 1,24p "$ROOT/tests/golden/vectors256.txt" > len16/vec.txt
(cd act && $R activation act.bin > log.txt 2>&1) &
for d in full1 v6full full2 full3; do (cd $d && $R eval vec.txt eval.bin > log.txt 2>&1) & done
(cd glue && $R eval vec.txt eval.bin --glue-only > log.txt 2>&1) &
(cd len16 && $R eval vec.txt eval.bin --length 16 > log.txt 2>&1) &
wait
echo "reference dumps are in $G; now run: python tests/golden/make_fixtures.py"
