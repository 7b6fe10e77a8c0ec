"""Summary of an ncu --set full report: python profiles/ncu_summary.py report.ncu-rep [traffic-key]
With a traffic key (e.g. realistic:8x256x50x64) the per-launch DRAM bytes are merged into profiles/traffic.json."""
import csv,subprocess,sys,io,json,os
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
hdr=rows[0]; units=rows[1]; data=rows[2:]
want=["Kernel Name","gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","sm__throughput.avg.pct_of_peak_sustained_elapsed","sm__warps_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","launch__occupancy_limit_registers","launch__occupancy_limit_shared_mem","launch__occupancy_limit_warps","launch__occupancy_limit_blocks","sm__inst_executed.sum","smsp__inst_executed.avg.per_cycle_active","sm__inst_executed_pipe_xu.sum","smsp__issue_active.avg.pct_of_peak_sustained_active","smsp__thread_inst_executed_per_inst_executed.ratio","launch__grid_size","launch__shared_mem_per_block_dynamic","smsp__warps_eligible.avg.per_cycle_active","l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum","lts__t_sectors_op_read.sum","smsp__cycles_active.avg","sm__cycles_elapsed.max","smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio","smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio"]
for d in data:
    print("="*100)
    for w in want:
        if w in hdr:
            i=hdr.index(w); print(f"{w:80s} {d[i]:>20s} {units[i]}")
    # stall reasons
    st=[(hdr[i],d[i]) for i in range(len(hdr)) if "smsp__average_warps_issue_stalled" in hdr[i] and "per_issue_active" in hdr[i]]
    st=sorted(st,key=lambda x:-float(x[1].replace(',','') or 0))[:8]
    for k,v in st: print("   stall",k.replace("smsp__average_warps_issue_stalled_","").replace("_per_issue_active.ratio",""),v)
    pipes=[(hdr[i],d[i]) for i in range(len(hdr)) if hdr[i].startswith("sm__inst_executed_pipe_") and hdr[i].endswith(".sum")]
    pipes=sorted(pipes,key=lambda x:-float(x[1].replace(',','') or 0))[:10]
    for k,v in pipes: print("   pipe",k,v)

if len(sys.argv)>2:
    key=sys.argv[2]; path=os.path.join(os.path.dirname(os.path.abspath(__file__)),"traffic.json")
    t=json.load(open(path)) if os.path.exists(path) else {}
    ent=t.setdefault(key,{})
    def num(d,name):
        i=hdr.index(name); v=float(d[i].replace(",","")); u=units[i].lower()
        return v*{"byte":1,"kbyte":1e3,"mbyte":1e6,"gbyte":1e9}[u]
    # a launch of pert_shade_fwd / pert_shade_bwd = main kernel + fallback kernel: sum both, per main launch
    tot={"pert_shade_fwd":0.0,"pert_shade_bwd":0.0}; n={"pert_shade_fwd":0,"pert_shade_bwd":0}
    for d in data:
        kn=d[hdr.index("Kernel Name")]
        name="pert_shade_fwd" if "shade_fwd" in kn else ("pert_shade_bwd" if "shade_bwd" in kn else None)
        if not name: continue
        tot[name]+=num(d,"dram__bytes_read.sum")+num(d,"dram__bytes_write.sum")
        if "fallback" not in kn: n[name]+=1
    for k in tot:
        if n[k]: ent[k]=tot[k]/n[k]
    # issue-slot utilisation and duration of every kernel of the step (the path is issue-bound, not HBM-bound)
    kern={}
    for d in data:
        kn=d[hdr.index("Kernel Name")]
        short=("fwd" if "shade_fwd" in kn else "bwd")+("_fallback" if "fallback" in kn else "_main")
        if "shade_fwd_fallback" in kn: short+="_coverage" if ", 16>" in kn else "_aggregate"
        i=hdr.index("gpu__time_duration.sum"); us=float(d[i].replace(",",""))*{"ns":1e-3,"us":1,"ms":1e3,"s":1e6}.get(units[i].lower(),1)
        kern[short]={"issue_active_pct":round(float(d[hdr.index("smsp__issue_active.avg.pct_of_peak_sustained_active")].replace(",","")),1),"ncu_us":round(us,1)}
    ent["kernels"]=kern
    ent["source"]=os.path.basename(rep)
    json.dump(t,open(path,"w"),indent=1)
