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
import re,json,sys
s=open(sys.argv[1]).read()
mul={'Gbyte':1e9,'Mbyte':1e6,'Kbyte':1e3,'byte':1}
m1=re.search(r'dram__bytes_read.sum,(\w+),([\d.]+)',s); m2=re.search(r'dram__bytes_write.sum,(\w+),([\d.]+)',s)
tot=float(m1.group(2))*mul[m1.group(1)]+float(m2.group(2))*mul[m2.group(1)]
json.dump({"world":"push","envs":65536,"horizon":2000,"bytes_per_launch":tot,"source":sys.argv[1].split('/')[-1]+" (dram__bytes_read.sum + dram__bytes_write.sum, one ncu --set full capture)"},open(sys.argv[2]+'/profiles/dram_traffic.json','w'),indent=1)
print(tot/1e9,'GB DRAM traffic per launch')
PY
