#!/bin/bash
# Round-end evidence on ONE GPU: full GPU test suite, smoke, the default bench line, the reference arm, the ncu launch
# list and the --set full capture of every kernel of one step (both from the SAME bench command, eager launches).
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
CMD="python bench.py --steps 4 --warmup 3 --no-graph --no-cpu-baseline --no-beside"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_l.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_fused|k_prep|k_chain|k_update' -s 15 -c 10 -f -o gpurun_out/${TAG}_full $CMD > gpurun_out/${TAG}_ncu_f.log 2>&1
timeout 300 python bench.py --workload wikikg2_rotate --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_rotate.json 2> gpurun_out/${TAG}_bench_rotate.err
# race check of the fused step (shared-memory hazards between the TMA / MMA / epilogue / prefetch roles) on one small case
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_fused.py -q -m gpu -x -k "five_launches and TransE_l2_d64" > gpurun_out/${TAG}_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|error" gpurun_out/${TAG}_racecheck.log | tail -4
ls -la gpurun_out/${TAG}_full.ncu-rep
python - <<P
import json
for n in ('bench','bench_ref','bench_rotate'):
    try:
        txt=open('gpurun_out/${TAG}_%s.json'%n).read(); d=json.loads(txt[txt.index('{'):])
        print(n,'value %.2fM e2e %.2fM ms %s'%(d['value']/1e6,d['e2e']['value']/1e6,d.get('ms_per_step')), 'frac', d.get('roofline',{}).get('frac'), 'cpu', d.get('cpu_baseline'))
        if 'beside' in d: print('  beside %.2fM'%(d['beside']['value']/1e6), d['beside']['config']['workload'][:40])
    except Exception as e: print(n,'ERR',e)
P
