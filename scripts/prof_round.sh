# usage: scripts/prof_round.sh TAG [bench args]   -- plain run, ncu launch list, ncu --set full of the tile sweep
TAG=${1:-r02}; shift
ARGS="--steps 3 --warmup 3 --no-cpu-baseline --no-e2e $*"
set -x
python bench.py $ARGS > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py $ARGS > gpurun_out/${TAG}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep_tile -s 3 -c 1 -f -o gpurun_out/prof_${TAG} python bench.py $ARGS > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out/prof_${TAG}.ncu-rep
