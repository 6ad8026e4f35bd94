"""Summarise build/csrc/*.ptxas.log: registers / spills / smem per kernel (demangled, short)."""
import glob, re, subprocess, sys
pat = sys.argv[1] if len(sys.argv) > 1 else ""
for f in sorted(glob.glob("/root/repo/build/csrc/*.ptxas.log")):
    txt = open(f).read()
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes smem)?", txt, re.S):
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"dpp::\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(.*", "", name).replace("void ", "")
        if pat and pat not in name: continue
        print(f"{name:55s} regs={m.group(5):>3s} stack={m.group(2):>4s} spill_st={m.group(3):>4s} smem={m.group(6) or 0}")
