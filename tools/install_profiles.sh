#!/bin/bash
# copy one tools/round_profiles.sh pass (gpurun_out/<TAG>_*) into profiles/r02_* and rebuild the summaries + traffic.json
T=${1:-r02w}; O=gpurun_out
cp $O/${T}_bench.json profiles/r02_bench.json
cp $O/${T}_bench_reference.json profiles/r02_bench_reference_arm.json
cp $O/${T}_launches.csv profiles/r02_ncu_launches.csv
cp $O/${T}_phase_profile.txt profiles/r02_phase_profile.txt
cp $O/${T}_phase_profile_synthetic.txt profiles/r02_phase_profile_synthetic.txt
cp $O/${T}_bundled_latency.jsonl profiles/r02_bundled_latency.jsonl
ncu -i $O/${T}_band.ncu-rep --page raw --csv 2>/dev/null > profiles/r02_ncu_band_metrics.csv
ncu -i $O/${T}_general.ncu-rep --page raw --csv 2>/dev/null > profiles/r02_ncu_general_wide_metrics.csv
python tools/ncu_lines.py $O/${T}_band.ncu-rep mcc_band_kernelILi512ELi1 30 > profiles/r02_ncu_band512_by_line.txt
NCU_KERNEL=unstru_kernel python tools/ncu_lines.py $O/${T}_band.ncu-rep unstru_kernel 20 > profiles/r02_ncu_unstru_by_line.txt
python tools/ncu_lines.py $O/${T}_general.ncu-rep mcc_persistentILi1ELi10 30 > profiles/r02_ncu_general_wide_by_line.txt
NCU_ARGS="--launch-count 1" python tools/hot_code.py $O/${T}_band.ncu-rep > profiles/r02_ncu_band512_hot_code.txt 2>&1
python - <<'PY'
import csv, json, subprocess
vals={}
for f in ['profiles/r02_ncu_band_metrics.csv','profiles/r02_ncu_general_wide_metrics.csv']:
    rows=list(csv.reader(open(f)))
    h=rows[0]; idx={n:i for i,n in enumerate(h)}; u=rows[1]
    for v in rows[2:]:
        g=lambda n: (float(v[idx[n]]), u[idx[n]])
        k=v[idx['Kernel Name']]
        vals[k]=(g('dram__bytes_read.sum'), g('dram__bytes_write.sum'), g('gpu__time_duration.sum'), g('lts__t_sector_hit_rate.pct'), g('smsp__issue_active.avg.pct_of_peak_sustained_active'), g('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'))
        print(k[:45], vals[k])
def B(x):
    v,u=x
    return v*{'Gbyte':1e9,'Tbyte':1e12,'Mbyte':1e6}[u]
head=subprocess.run(['git','rev-parse','--short','HEAD'],stdout=subprocess.PIPE,text=True).stdout.strip()
b512=[k for k in vals if '512, 1' in k][0]; b256=[k for k in vals if '256, 2' in k][0]; un=[k for k in vals if 'unstru' in k][0]; gen=[k for k in vals if 'mcc_persistent' in k][0]
t={
 "mica_ompa": {
  "bytes_per_launch": B(vals[b512][0])+B(vals[b512][1]),
  "ncu_report": f"profiles/r02_ncu_band_metrics.csv (ncu --set full, 1000 MicA x ompA shuffles; row mcc_band_kernel<512,1>: dram__bytes_read.sum {B(vals[b512][0])/1e9:.3f} GB + dram__bytes_write.sum {B(vals[b512][1])/1e9:.3f} GB)",
  "git_head_of_capture": head+" (end of round 2)",
  "note": f"the dominant kernel only: 2000 of the 3000 problems of a step.  Same capture: mcc_band_kernel<256,2> {B(vals[b256][0])/1e9:.3f} + {B(vals[b256][1])/1e9:.3f} GB, unstru_kernel {B(vals[un][0])/1e9:.2f} + {B(vals[un][1])/1e9:.2f} GB"
 },
 "synthetic": {
  "bytes_per_launch": B(vals[gen][0])+B(vals[gen][1]),
  "pairs_per_launch": 148,
  "ncu_report": f"profiles/r02_ncu_general_wide_metrics.csv (ncu --set full of mcc_persistent<1,10>, 148 pairs of the 1000 x 500 batch: dram__bytes_read.sum {B(vals[gen][0])/1e12:.3f} TB + dram__bytes_write.sum {B(vals[gen][1])/1e9:.1f} GB = {(B(vals[gen][0])+B(vals[gen][1]))/148/1e9:.1f} GB per pair)",
  "git_head_of_capture": head+" (end of round 2)"
 }
}
json.dump(t,open('profiles/traffic.json','w'),indent=1)
j=json.loads(open('profiles/r02_bench.json').read().strip().splitlines()[-1])
r=j['roofline']
print('value',j['value'],j['ms_per_step'],'e2e',j['e2e']['value'],j['e2e_sparse']['value'],'frac',r['frac'],r['launch_ms'],'whole',r['whole_step']['frac'],r['whole_step']['ms'],'cpu',j['cpu_baseline']['value'],j['cpu_baseline']['single_thread_value'])
c=j['config4']; print('c4',c['value'],c['ms_per_step'],c['e2e']['value'],c['roofline']['frac'],c['roofline']['launch_ms'],c['cpu_baseline']['value'])
r=json.loads(open('profiles/r02_bench_reference_arm.json').read().strip().splitlines()[-1]); print('ref',r['value'],r['config4']['value'])
PY
