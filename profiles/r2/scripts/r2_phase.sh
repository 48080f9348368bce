# K2's SM clocks per phase (build: make -C scbonita_b200/csrc EXTRA=-DSCB_PHASE_CLOCKS OUT=../libscbonita_phase.so OBJDIR=_obj_phase)
SCB_B200_LIB=scbonita_b200/libscbonita_phase.so python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-cases --opt graph=0 --opt fuse=0 > gpurun_out/phase.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/phase.json')); s=d['stats']
tot=s['cyc_prologue']+s['cyc_steps']+s['cyc_epilogue']
print('ms',d['ms_per_step'],'k2',s['sim_ms'],{k:(s[k],round(s[k]/tot,3)) for k in ('cyc_prologue','cyc_steps','cyc_epilogue')},'tiles',s['tiles'],'steps',s['steps_executed'])
PY
