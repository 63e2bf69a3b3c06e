import re,subprocess,sys
pat=sys.argv[2] if len(sys.argv)>2 else ''
txt=open(sys.argv[1]).read()
blocks=re.split(r"ptxas info\s+: Compiling entry function '", txt)[1:]
for b in blocks:
    name=b.split("'")[0]
    dem=subprocess.run(['c++filt',name],capture_output=True,text=True).stdout.strip()
    if pat in dem:
        m=re.search(r"Used (\d+) registers",b); sp=re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads",b); sm=re.search(r"(\d+) bytes smem",b)
        print(dem[:100], 'regs',m.group(1),'spill',sp.groups() if sp else None,'smem',sm.group(1) if sm else 0)
