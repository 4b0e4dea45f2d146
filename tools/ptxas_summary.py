"""Summarise `-Xptxas -v` logs written by model_predictive_control_b200/_build.py."""
import glob, os, re, subprocess, sys
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "model_predictive_control_b200", "build")
pat = sys.argv[1] if len(sys.argv) > 1 else ""
for log in sorted(glob.glob(os.path.join(root, "*.ptxas.log"))):
    name = None
    for line in open(log):
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("void mpc::", "")
            spill = ""
        if "bytes stack frame" in line:
            spill = line.strip()
        m = re.search(r"Used (\d+) registers(.*)", line)
        if m and name and pat in name:
            sm = re.search(r"(\d+) bytes smem", line)
            sp = re.findall(r"(\d+) bytes", spill)
            print(f"{name:70s} regs={m.group(1):>3s} stack/spill_st/spill_ld={'/'.join(sp)}")
