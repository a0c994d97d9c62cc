import json,sys,subprocess,os
out=subprocess.run([sys.executable,"bench.py","--steps","10","--no-cpu","--no-e2e","--no-probe"],capture_output=True,text=True).stdout
d=json.loads(out.strip().splitlines()[-1]); g=d["griffin_lim"]
print(os.environ.get("SPEECHDSP_LIB","in-tree"),"GL ms/iter", g["ms_per_iteration"], "frac", g["roofline"]["frac"])
