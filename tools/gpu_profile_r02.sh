set -u
CMD="python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline --no-parity --rows-per-gpu 2"
$CMD > gpurun_out/r02r_plain.json 2> gpurun_out/r02r_plain.err || { echo plain failed; tail -5 gpurun_out/r02r_plain.err; exit 1; }
python -c "import json;d=json.loads(open('gpurun_out/r02r_plain.json').read().strip().splitlines()[-1]);print('value',d['value'],'e2e',d['e2e']['value'],d['net_stage_ms_first_chunk'])"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02r_launches.csv $CMD > gpurun_out/r02r_ncu_launches.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:fused_block_kernel -c 1 -f -o gpurun_out/r02r_block1 $CMD > gpurun_out/r02r_ncu_block1.log 2>&1; echo "block1 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:sep_uf_kernel -c 2 -f -o gpurun_out/r02r_sepuf $CMD > gpurun_out/r02r_ncu_sepuf.log 2>&1; echo "sepuf rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"stft_db_kernel|select_hist|pool_res_f32|conv0_direct" -c 6 -f -o gpurun_out/r02r_misc $CMD > gpurun_out/r02r_ncu_misc.log 2>&1; echo "misc rc=$?"
ls -la gpurun_out/r02r_*.ncu-rep
