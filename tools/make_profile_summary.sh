#!/bin/bash
# usage: tools/make_profile_summary.sh <report.ncu-rep> <launches.csv> <tag>   (run right after the capture, same libprl.so)
set -e
REP=$1; LAUNCH=$2; TAG=$3
ROOT=$(cd "$(dirname "$0")/.." && pwd)
cp $LAUNCH $ROOT/profiles/${TAG}_launches.csv
OUT=$ROOT/profiles/${TAG}_rollout_summary.md
{
echo "# ${TAG}: \`rollout_kernel<2, DevWorldSmall, 448>\` on push.json, 65536 envs x 2000 steps (BASELINE configs[2])"
echo
echo "Captured with \`ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 3 -c 1\` after the same"
echo "\`python bench.py --steps 2 --warmup 3 --no-cpu-baseline\` exited 0 without ncu. Launch list of the same command:"
echo "\`profiles/${TAG}_launches.csv\` (the rollout kernel is >99.9 % of a step; the other launches are the 70 us L2-flush fills)."
echo
echo "Key metrics, stall mix and the hottest source lines (tools/ncu_summary.sh):"
echo; echo '```'
$ROOT/tools/ncu_summary.sh $REP rollout_kernelILi2 30
echo '```'; echo
echo "By function (tools/ncu_regions.py):"; echo; echo '```'
TMP=$(mktemp -d); (cd $TMP && cuobjdump -xelf all $ROOT/physics_reinforcement_learning_environment_b200/libprl.so >/dev/null)
ncu -i $REP --page source --csv 2>/dev/null > $TMP/src.csv
python $ROOT/tools/ncu_regions.py $TMP/src.csv rollout_kernelILi2 $TMP/prl_api.sm_100a.cubin $ROOT/physics_reinforcement_learning_environment_b200/csrc 2>/dev/null | head -32
echo '```'
} > $OUT
python - "$OUT" "$ROOT" <<'PY'
# the ncu figures bench.py reports in `roofline` (traffic, issue), keyed by the kernel source hash
import re,json,sys
sys.path.insert(0, sys.argv[2])
import bench
s=open(sys.argv[1]).read()
mul={'Gbyte':1e9,'Mbyte':1e6,'Kbyte':1e3,'byte':1}
m1=re.search(r'dram__bytes_read.sum,(\w+),([\d.]+)',s); m2=re.search(r'dram__bytes_write.sum,(\w+),([\d.]+)',s)
tot=float(m1.group(2))*mul[m1.group(1)]+float(m2.group(2))*mul[m2.group(1)]
val=lambda name: float(re.search(name+r',[^,]*,([\d.]+)',s).group(1))
rec={"world":"push","envs":65536,"horizon":2000,"source_sha256":bench.kernel_source_hash(),
     "dram_bytes_per_launch":tot,
     "warp_instructions_per_launch":val('smsp__inst_executed.sum'),
     "issue_active_frac":val('smsp__issue_active.avg.pct_of_peak_sustained_active')/100.0,
     "active_lanes_per_inst":val('smsp__thread_inst_executed_per_inst_executed.ratio'),
     "kernel_ms_under_ncu":val('gpu__time_duration.sum'),
     "source":"profiles/"+sys.argv[1].split('/')[-1]+" (one ncu --set full capture of the rollout kernel)"}
json.dump(rec,open(sys.argv[2]+'/profiles/rollout_profile.json','w'),indent=1)
print(rec)
PY
