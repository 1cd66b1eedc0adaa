#!/usr/bin/env bash
# A/B of training options on nerf/fox: PSNR@1000 on the held-out views, ms/step, counters.  usage: tools/psnr_ab.sh "name=value ..." [repeats]
O=gpurun_out/psnr_ab; mkdir -p $O
OPTS=""; for kv in $1; do OPTS="$OPTS --option $kv"; done
TAG=$(echo "$1" | tr ' =.' '___'); N=${2:-2}
for i in $(seq 1 $N); do
  timeout 300 python tools/ref_app.py --impl ngp_b200 --scene fox --enc L16F2 --steps 1000 --render-repeats 1 $OPTS --out $O/${TAG:-default}_$i.json 2>&1 | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('$1', '#$i', 'psnr %.4f' % j['psnr_mean'], 'ms %.3f' % j['ms_per_step'], j['counters'])"
  rm -f $O/*.ingp $O/*.npz
done
