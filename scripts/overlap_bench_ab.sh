for ov in 0 1 0 1; do
  TRAY_CUDA_OVERLAP=$ov TRAY_CUDA_GEN_MIN=4 timeout 200 python bench.py --steps 100 2>/dev/null > /tmp/ob.json
  python -c "
import json; b=json.load(open('/tmp/ob.json')); print('overlap $ov', round(b['value'],1), round(b['ms_per_step'],4), 'e2e', round(b['e2e']['value'],1), 'warm', round(b['warm_l2']['value'],1))"
done
