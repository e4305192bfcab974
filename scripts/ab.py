import sys, os, shutil, subprocess, json
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, 'vegasafterglow_b200', 'libvag_b200.so')
for name in sys.argv[1:]:
    alt = os.path.join(root, 'vegasafterglow_b200', name)
    shutil.copy(lib, lib + '.bak'); shutil.copy(alt, lib)
    out = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--steps', '10', '--warmup', '3', '--no-cpu-baseline'], capture_output=True, text=True).stdout.strip().splitlines()[-1]
    d = json.loads(out); print(name, d['value'], d['roofline_fp64']['ms_per_step'], d['loglike']['value'], flush=True)
    shutil.copy(lib + '.bak', lib)
