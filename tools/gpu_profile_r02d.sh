# Final-state profile set of round 2 (run on the GPU box through gpurun; outputs under gpurun_out/, summaries copied to profiles/):
#   plain bench line -> launch list of the same command -> ONE `ncu --set full` run over the first predict call's hot kernels,
#   condensed on the box (tools/ncu_summary.py) because the .ncu-rep files exceed what gpurun brings back.
set -u
TAG=${1:-r02zv}
CMD="python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline --no-parity --rows-per-gpu 2"
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { echo plain failed; tail -5 gpurun_out/${TAG}_plain.err; exit 1; }
python -c "import json;d=json.loads(open('gpurun_out/${TAG}_plain.json').read().strip().splitlines()[-1]);print('value',d['value'],'e2e',d['e2e']['value'],d['stage_ms'],d['net_stage_ms_first_chunk'])"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "launch list rc=$?"
[ "${SKIP_FULL:-0}" = "1" ] && exit 0
ncu --set full --clock-control none -k regex:"fused_block_kernel|sep_uf_kernel|pool_res_f32|conv0_direct|lstm_rec|gemm_tc_kernel|stft_db|select_hist" -c 34 -f -o /tmp/${TAG}_full $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "full rc=$?"
python tools/ncu_summary.py /tmp/${TAG}_full.ncu-rep --out gpurun_out/${TAG}_ncu_full_summary.csv
ls -la /tmp/${TAG}_full.ncu-rep gpurun_out/
