"""run single UMMA self-test cases in separate processes (a faulting case must not poison the rest)"""
import subprocess, sys
cases = sys.argv[1:] or ["test_umma_mixed_f16_bf16_operands", "test_umma_bf16_mn_major_b_half_swizzle_row", "test_umma_bf16_mn_major_b_second_half_of_the_row"]
ids = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_umma.py", "-m", "gpu", "--collect-only", "-q"], capture_output=True, text=True).stdout.split("\n")
for nid in ids:
    if "::" not in nid or not any(c in nid for c in cases):
        continue
    r = subprocess.run([sys.executable, "-m", "pytest", nid, "-m", "gpu", "-q", "--timeout", "60", "-x", "--tb=line"], capture_output=True, text=True)
    tail = [l for l in r.stdout.split("\n") if l.strip()][-3:]
    print(nid.split("::")[1], "->", " | ".join(t[:150] for t in tail))
