set -u
CMD="python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline --no-parity --rows-per-gpu 2"
$CMD > gpurun_out/r02zn_plain.json 2> gpurun_out/r02zn_plain.err || { echo plain failed; tail -5 gpurun_out/r02zn_plain.err; exit 1; }
python -c "import json;d=json.loads(open('gpurun_out/r02zn_plain.json').read().strip().splitlines()[-1]);print('value',d['value'],'e2e',d['e2e']['value'],d['net_stage_ms_first_chunk'])"
ncu --set full --clock-control none --import-source on -k regex:"pool_res_f32|gemm_tc_kernel|lstm_rec_kernel|conv0_direct" -c 8 -f -o gpurun_out/r02zn_misc $CMD > gpurun_out/r02zn_ncu_misc.log 2>&1; echo "misc rc=$?"
ls -la gpurun_out/r02zn_*.ncu-rep
