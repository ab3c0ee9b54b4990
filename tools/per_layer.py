#!/usr/bin/env python
"""Map the GEMM launches of one training step (ncu launch list of tools/profile_step.py, B=64 224x224) to layers and
print achieved TFLOP/s per launch (developer tool)."""
import csv, re, sys
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
S = int(sys.argv[3]) if len(sys.argv) > 3 else 224
lines = [l for l in open(path) if not l.startswith("==")]
gem = []
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum" or "gemm" not in row["Kernel Name"]:
        continue
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    gem.append((re.sub(r"\(.*", "", row["Kernel Name"]).replace("void cs::", ""), v))
LC = {1: 64, 2: 128, 3: 256, 4: 512, 5: 1024}; LH = {L: S >> (L - 1) for L in range(1, 6)}
convs = []
for i in range(18):
    if i < 10:
        L = i // 2 + 1; cout = LC[L]; cin = LC[L] if i % 2 else (64 if L == 1 else LC[L - 1])
    else:
        L = 4 - (i - 10) // 2; cout = LC[L]; cin = LC[L] if i % 2 else 2 * LC[L]
    convs.append((L, cin, cout))
fc = lambda i: 2 * B * LH[convs[i][0]] ** 2 * convs[i][1] * convs[i][2] * (1 if i == 0 else 9)
fu = lambda k: 2 * B * LH[5 - k] ** 2 * LC[5 - k] * LC[4 - k] * 4
# Launches are matched per kernel family, in the order each family is issued (the two backward streams interleave in the
# ncu list, so a single global order does not exist):
#   conv family (stem / conv3_gemm / pix_gemm2 with CARTSEG_CONV3=0): fprop conv0..9, then per decoder level up-conv is
#       in the pix family; dgrads conv17..1 in backward order
#   pix family: conv-transpose fprop up0..3, then their dgrads up3..0
#   wgrad family (side stream): conv17, conv16, up3, conv15, conv14, up2, conv13, conv12, up1, conv11, conv10, up0, conv9..0
fam = lambda kn: "wgrad" if "wgrad" in kn else ("pix" if kn.startswith("pix_gemm2") else "conv")
seq = {"conv": [("fprop", f"conv{i}", fc(i)) for i in range(10)], "pix": [], "wgrad": []}
for k in range(4):
    seq["pix"].append(("fprop", f"up{k}", fu(k)))
    seq["conv"] += [("fprop", f"conv{10+2*k}", fc(10 + 2 * k)), ("fprop", f"conv{11+2*k}", fc(11 + 2 * k))]
for i in range(17, 0, -1):
    seq["conv"].append(("dgrad", f"conv{i}", fc(i)))
for k in (3, 2, 1, 0):
    seq["pix"].append(("dgrad", f"up{k}", fu(k)))
    seq["wgrad"] += [("wgrad", f"conv{11+2*k}", fc(11 + 2 * k)), ("wgrad", f"conv{10+2*k}", fc(10 + 2 * k)), ("wgrad", f"up{k}", fu(k))]
seq["wgrad"] += [("wgrad", f"conv{i}", fc(i)) for i in range(9, -1, -1)]
pos = {k: 0 for k in seq}
order = []
for kn, us in gem:
    f = fam(kn)
    assert pos[f] < len(seq[f]), (f, kn)
    order.append(seq[f][pos[f]]); pos[f] += 1
assert all(pos[f] == len(seq[f]) for f in seq), pos
tot_t = tot_f = 0
for (kind, name, fl), (kn, us) in zip(order, gem):
    if name.startswith("conv"):
        L, cin, cout = convs[int(name[4:])]; lay = f"{name} L{L} {cin}->{cout}"
    else:
        k = int(name[2:]); lay = f"{name} L{5-k}->L{4-k} {LC[5-k]}->{LC[4-k]}"
    print(f"{kind:6s} {lay:26s} {us:8.1f} us {fl/us/1e6:7.1f} TF/s  {kn}")
    tot_t += us; tot_f += fl
print(f"all GEMMs: {tot_t:.0f} us, {tot_f/1e12:.2f} TFLOP, {tot_f/tot_t/1e6:.0f} TF/s")
