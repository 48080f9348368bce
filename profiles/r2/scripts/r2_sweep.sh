# option sweep: one bench line per variant (arg 1 = tag, rest = variants, each a quoted string of --opt flags)
mkdir -p gpurun_out
T=$1; shift
i=0
for V in "$@"; do
  i=$((i+1))
  timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-cases $V > gpurun_out/${T}_s$i.json 2> gpurun_out/${T}_s$i.err; rc=$?
  python - "$T" "$i" "$V" "$rc" <<'PY'
import json,sys
T,i,V,rc=sys.argv[1:5]
try:
    d=json.load(open('gpurun_out/%s_s%s.json'%(T,i)))
    st=d['stats']
    print('[%s]'%V,'ms',round(d['ms_per_step'],3),'k2',round(d['roofline']['kernel_ms'],3),'smem',round(d['roofline_binding']['frac'],3),'rows',st['state_rows_moved'],'steps',st['steps_executed'],'tiles',st['tiles'],'res_chunks',st['resolver_chunks'],'chk',d['fitness_checksum'],'geom',d['config']['geometry']['words_per_lane'],d['config']['geometry']['warps'],d['config']['geometry']['ctas'])
except Exception as e:
    print('[%s] rc=%s failed: %s'%(V,rc,e)); print(open('gpurun_out/%s_s%s.err'%(T,i)).read()[-600:])
PY
done
