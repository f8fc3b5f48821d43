set -u
CMD="python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline --no-parity --rows-per-gpu 2"
$CMD > gpurun_out/r02zp_plain.json 2> gpurun_out/r02zp_plain.err || { echo plain failed; tail -5 gpurun_out/r02zp_plain.err; exit 1; }
python -c "import json;d=json.loads(open('gpurun_out/r02zp_plain.json').read().strip().splitlines()[-1]);print('value',d['value'],'e2e',d['e2e']['value'],d['net_stage_ms_first_chunk'])"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02zp_launches.csv $CMD > gpurun_out/r02zp_ncu_launches.log 2>&1; echo "launch list rc=$?"
for K in conv0_direct fused_block_kernel sep_uf_kernel pool_res_f32 lstm_rec_kernel gemm_tc_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -c 2 -f -o gpurun_out/r02zp_$K $CMD > gpurun_out/r02zp_ncu_$K.log 2>&1; echo "$K rc=$?"
done
ls -la gpurun_out/r02zp_*.ncu-rep
