mkdir -p gpurun_out
R=$PWD
prof() { # name lib opts
  name=$1; lib=$2; shift; shift
  CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --opt fuse=0 $@"
  SCB_B200_LIB=$lib timeout 300 $CMD > gpurun_out/ab_plain_$name.log 2>&1 && \
  SCB_B200_LIB=$lib timeout 600 ncu --set full --clock-control none --import-source on -k regex:scb_k_sim -s 1 -c 1 -f -o gpurun_out/ab_$name $CMD > gpurun_out/ab_ncu_$name.log 2>&1
  echo "$name rc=$?"
}
prof old $R/gpurun_variants/lib_old.so
prof cur_s0 $R/scbonita_b200/libscbonita_b200.so --opt steady=0
prof cur_s1 $R/scbonita_b200/libscbonita_b200.so
ls -la gpurun_out/ab_*.ncu-rep
