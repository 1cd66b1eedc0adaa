#!/usr/bin/env bash
# A/B of training options on nerf/fox: PSNR@1000 on the held-out views, ms/step, counters.  usage: tools/psnr_ab.sh "name=value ..." [repeats]
O=gpurun_out/psnr_ab; mkdir -p $O
OPTS=""; for kv in $1; do OPTS="$OPTS --option $kv"; done
TAG=$(echo "$1$EXTRA" | tr ' =.-' '____'); N=${2:-2}
for i in $(seq 1 $N); do
  timeout 300 python tools/ref_app.py --impl ngp_b200 --scene fox --enc L16F2 --steps 1000 --render-repeats 1 $EXTRA $OPTS --out $O/${TAG:-default}_$i.json > $O/last.log 2>&1; python -c "
import json,sys
j=json.load(open('$O/${TAG:-default}_$i.json')); print('$1', '#$i', 'psnr %.4f' % j['psnr_mean'], 'ms %.3f' % j['ms_per_step'], 'views 4/8: %.2f %.2f' % (j['psnr_per_view'][4], j['psnr_per_view'][8]), 'loss %.5f' % (sum(v for k, v in j['loss_curve'].items() if int(k) > 800) / max(1, sum(1 for k in j['loss_curve'] if int(k) > 800))), j['counters'])"
  rm -f $O/*.ingp; [ -n "$KEEP_NPZ" ] || rm -f $O/*.npz
done
