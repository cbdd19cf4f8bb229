# usage: _ncu_one.sh <out-name> <kernel-regex> <count> <quick_time args...>
out=$1; shift; rx=$1; shift; cnt=$1; shift
python tools/quick_time.py --pairs 1 --reps 1 "$@" > gpurun_out/plain_$out.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"$rx" -c $cnt -o gpurun_out/$out python tools/quick_time.py --pairs 1 --reps 1 "$@" > gpurun_out/ncu_$out.log 2>&1
ls -la gpurun_out/$out.ncu-rep
