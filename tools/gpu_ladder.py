"""Runs a ladder of isolated GPU parity cases (smallest first); each in its own process with a timeout."""
import subprocess
import sys
import os

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [
    ["--n", "200", "--m", "512", "--cg", "1"],
    ["--n", "200", "--m", "512", "--cg", "2"],
    ["--n", "300", "--m", "1000", "--cg", "1", "--chunk", "256"],
    ["--n", "300", "--m", "1000", "--cg", "2", "--chunk", "256"],
    ["--n", "777", "--m", "3000", "--cg", "2", "--operand", "bf16"],
    ["--n", "777", "--m", "3000", "--cg", "2", "--operand", "fp16x2", "--chunk", "1024"],
    ["--n", "777", "--m", "3000", "--cg", "1", "--count", "popc", "--chunk", "1024"],
    ["--n", "500", "--m", "2000", "--cg", "2", "--flat"],
    ["--n", "600", "--m", "3000", "--cg", "2", "--chunk", "256", "--resident", "1"],
    ["--n", "600", "--m", "3000", "--cg", "2", "--chunk", "256", "--resident", "4"],
    ["--n", "600", "--m", "3000", "--cg", "1", "--chunk", "256", "--resident", "5", "--operand", "fp16x2"],
    ["--n", "300", "--m", "1000", "--cg", "1", "--count", "fp4"],
    ["--n", "300", "--m", "1000", "--cg", "2", "--count", "fp4"],
    ["--n", "777", "--m", "3000", "--cg", "2", "--count", "fp4", "--chunk", "1024", "--missing", "0.2"],
    ["--n", "2000", "--m", "10000", "--cg", "1"],
    ["--n", "2000", "--m", "10000", "--cg", "2"],
    ["--n", "2000", "--m", "10000", "--cg", "2", "--chunk", "2048"],
    ["--n", "2000", "--m", "10000", "--cg", "2", "--operand", "fp16x2"],
]


def main():
    extra = sys.argv[1:]
    results = []
    for c in CASES:
        cmd = [sys.executable, os.path.join(HERE, "gpu_case.py")] + c + extra
        print("=== case", " ".join(c), flush=True)
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
            print(r.stdout[-6000:])
            if r.returncode != 0:
                print("  [stderr tail]", r.stderr[-3000:])
            results.append((c, r.returncode))
        except subprocess.TimeoutExpired as ex:
            print("  TIMEOUT", (ex.stdout or b"")[-2000:], (ex.stderr or b"")[-2000:])
            results.append((c, "timeout"))
        sys.stdout.flush()
    print("=== summary")
    for c, rc in results:
        print(("PASS" if rc == 0 else f"FAIL({rc})"), " ".join(c))


if __name__ == "__main__":
    main()
