mkdir -p gpurun_out
run() { # name lib opts...
  name=$1; lib=$2; shift; shift
  SCB_B200_LIB=$lib timeout 300 python bench.py --no-cpu-baseline --steps 5 --warmup 3 "$@" > "gpurun_out/bis_$name.json" 2>> gpurun_out/bis.err
  python - "gpurun_out/bis_$name.json" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], 'ms', round(d['ms_per_step'],3), 'k2', round(d['roofline']['kernel_ms'],3), 'steps', d['stats']['steps_executed'], 'chk', d['fitness_checksum'])
PY
}
R=$PWD
run old $R/gpurun_variants/lib_old.so
run cur_s0 $R/scbonita_b200/libscbonita_b200.so --opt steady=0
run cur_s1 $R/scbonita_b200/libscbonita_b200.so
run norows_s0 $R/gpurun_variants/lib_norows.so --opt steady=0
run norows_s1 $R/gpurun_variants/lib_norows.so
run layout_s0 $R/gpurun_variants/lib_layout.so --opt steady=0
run layout_s1 $R/gpurun_variants/lib_layout.so
run old_again $R/gpurun_variants/lib_old.so
