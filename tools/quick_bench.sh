for lib in detectinblur_b200/libdib*.so; do   # drop experimental builds next to libdib.so to compare them
  echo "== $lib"
  for w in cfg2 cfg3 cfg2h; do
    DIB_LIB_PATH=$PWD/$lib timeout 120 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$w', round(d['value']), 'img/s', round(d['ms_per_step']*1000,1), 'us  hbm_frac', round(r['frac'],3), 'max_roof_frac', round(r['frac_of_max_roofline'],3))"
  done
done
