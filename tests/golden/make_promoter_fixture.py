"""Makes tests/golden/promoters_head.fa: the first 64 of the 500 promoter sequences the reference ships as cookbook data
(/root/reference/supplementary/cookbook/recipes/promoters.fa; DeNovoSunflower-style input).  Data only, unchanged; the GPU
box has no /root/reference, so the tests read this copy."""
import os

SRC = "/root/reference/supplementary/cookbook/recipes/promoters.fa"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "promoters_head.fa")

if __name__ == "__main__":
    out, n = [], 0
    for line in open(SRC):
        if line.startswith(">"):
            n += 1
            if n > 64:
                break
        out.append(line)
    open(DST, "w").write("".join(out))
    print(f"{DST}: {n - 1 if n > 64 else n} sequences, {sum(len(x) for x in out)} bytes")
