set -u
TAG=$1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 200 python scripts/ncu_target.py c5 1250000 5 4 --no-filter 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv python scripts/ncu_target.py c5 1250000 5 3 --no-filter > /dev/null 2>&1
