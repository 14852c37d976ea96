"""Developer tool: list the global-memory instructions of one kernel with their scheduling control fields (stall count,
write/read scoreboard, wait mask) and, for every load or atomic, the first later instruction that waits on its
scoreboard -- i.e. where its latency is actually exposed.  ptxas has six scoreboards per warp and decides which
loads share one; two independent loads on one scoreboard are waited for together.

    python scripts/sass_ctrl.py frontx_b200/libfrontx_b200.so _ZN2fx19inverse_scan_kernelILb0ELb1EEE

The control fields sit in bits 105..125 of the 128-bit instruction word (stall 4 bits, yield 1, write barrier 3, read
barrier 3, wait mask 6, reuse 4), the layout in use since sm_70; decoded from `nvdisasm -hex`.
"""
import re
import subprocess
import sys
import tempfile
from pathlib import Path


def kernel_sass(lib: str, kern: str):
    tmp = Path(tempfile.mkdtemp())
    subprocess.run(["cuobjdump", "-xelf", "all", str(Path(lib).resolve())], cwd=tmp, check=True, capture_output=True)
    cubin = next(tmp.glob("*.cubin"))
    lines = subprocess.run(["nvdisasm", "-c", "-hex", str(cubin)], capture_output=True, text=True).stdout.split("\n")
    start = [k for k, l in enumerate(lines) if l.startswith(".text.") and kern in l and l.rstrip().endswith(":")][0]
    out, k = [], start + 1
    while k < len(lines) and not lines[k].startswith(".text."):
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/", lines[k])
        m2 = re.match(r"\s*/\* (0x[0-9a-f]{16}) \*/", lines[k + 1]) if m and k + 1 < len(lines) else None
        if m and m2:
            hw = int(m2.group(1), 16)
            out.append({"addr": int(m.group(1), 16), "txt": m.group(2).strip(), "stall": (hw >> 41) & 0xF,
                        "wbar": (hw >> 46) & 7, "rbar": (hw >> 49) & 7, "wait": (hw >> 52) & 0x3F})
            k += 2
        else:
            k += 1
    return out


def main():
    ins = kernel_sass(sys.argv[1], sys.argv[2])
    for k, i in enumerate(ins):
        t = i["txt"]
        if not any(op in t for op in ("LDG.", "ATOMG", "LDGSTS", "DEPBAR", "STG.")):
            continue
        waits = "".join(str(b) for b in range(6) if (i["wait"] >> b) & 1)
        line = (f"{i['addr']:6x} stall {i['stall']:2d} W{i['wbar'] if i['wbar'] != 7 else '-'} "
                f"R{i['rbar'] if i['rbar'] != 7 else '-'} wait[{waits:6s}] {t[:84]}")
        if i["wbar"] != 7 and ("LDG." in t or "ATOMG" in t):
            for j in range(k + 1, len(ins)):
                if (ins[j]["wait"] >> i["wbar"]) & 1:
                    line += f"  -> first wait +{j - k} instr at {ins[j]['addr']:x}: {ins[j]['txt'][:36]}"
                    break
        print(line)


if __name__ == "__main__":
    main()
