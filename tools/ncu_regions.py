import csv, subprocess, collections, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr=None
fname=""
data=[]
for r in rows:
    if len(r)==2 and r[0]=="File Path": fname=r[1].split('/')[-1]; continue
    if len(r) > 10 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    d={h:r[i] for i,h in enumerate(hdr)}
    if not d['Line No'].isdigit(): continue
    try:
        data.append((fname, int(d['Line No']), int(d['# Samples']), int(d['Instructions Executed']), {k:int(v) for k,v in d.items() if k.startswith('stall_') and 'Not Issued' not in k and v.isdigit()}))
    except (ValueError, KeyError): continue
tot=sum(d[2] for d in data); toti=sum(d[3] for d in data)
print("total samples", tot, "instr", toti)
src=open('staarpipeline_b200/csrc/fused_tc.cu').read().split('\n')
def find(s): 
    for i,l in enumerate(src):
        if s in l: return i+1
marks=[("prologue", 1), ("expander", find("expanders: thread = variant row")), ("epilogue", find("epilogue: 7 digit sums")), ("scan-setup", find("// ===================================================================== scanners")),
 ("kb-loop-head", find("const int wi = (int)kb * 16 + l16;")), ("guess", find("orientation guess: more hom-2")), ("phase1", find("---- phase 1")), ("consume", find("look-ups issued in the previous k-block")), ("2a-missing", find("---- phase 2a")), ("lookups-fn", find("---- phase 2b")), ("push", find("const int par3 = (int)(kb % 3);")), ("barrier", find("list complete, pending look-ups consumed")), ("batches", find("const int total = min(")), ("tail", find("---- per-row totals")), ("mma", find("MMA issuer (one thread)")), ("digit-prod", find("digit producer (one thread)")), ("geno-prod", find("genotype producer (one thread)")), ("end", find("// Finaliser of the fused path"))]
agg=collections.OrderedDict((m[0],[0,0,collections.Counter()]) for m in marks)
for f,ln,smp,ins,st in data:
    if f!='fused_tc.cu':
        a=agg.setdefault(f,[0,0,collections.Counter()])
    else:
        name="prologue"
        for m,start in marks:
            if start and ln>=start: name=m
        a=agg[name]
    a[0]+=smp; a[1]+=ins; a[2].update(st)
for k,(smp,ins,st) in agg.items():
    print(f"{k:14s} samples {smp:7d} {smp/tot*100:5.1f}%  instr {ins:>10d} {ins/toti*100:5.1f}%  {[(a.replace('stall_',''),b) for a,b in st.most_common(5)]}")
