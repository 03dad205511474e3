"""Developer tool: analytic inventory of the yv_gemm launches of one cfg2 training step (shapes, tile counts) and the
tensor-pipe time they need at full SM fill vs with wave quantisation.  No GPU needed."""
import math
P = 8; Mv = P * 288; Mt = P * 80
CYC_PER_K16 = 3 * 64          # bf16x3: three 128x128x16 MMAs, 64 cycles each
CLK = 1.9e9; SMS = 148
L = []                        # (name, M, N, K, batch, count)
def lin(name, M, N, K, cnt, dgrad=True):
    L.append((name + " fwd", M, N, K, 1, cnt))
    if dgrad: L.append((name + " dgrad", M, K, N, 1, cnt))
    L.append((name + " wgrad", N, K, M, 1, cnt))
def attn(name, pairs, heads, Tq, Tk, dh, cnt):
    B = pairs * heads
    L.append((name + " QK^T", Tq, Tk, dh, B, cnt)); L.append((name + " PV", Tq, dh, Tk, B, cnt))
    L.append((name + " dP", Tq, Tk, dh, B, cnt)); L.append((name + " dV", Tk, dh, Tq, B, cnt))
    L.append((name + " dQ", Tq, dh, Tk, B, cnt)); L.append((name + " dK", Tk, dh, Tq, B, cnt))
lin("img embed", Mv, 1024, 2048, 1, dgrad=False)
lin("text qkv", Mt, 2304, 768, 12); attn("text attn", P, 12, 80, 80, 64, 12); lin("text out", Mt, 768, 768, 12)
lin("text ffn1", Mt, 3072, 768, 18); lin("text ffn2", Mt, 768, 3072, 18)
lin("vis qkv", Mv, 3072, 1024, 6); attn("vis attn", P, 8, 288, 288, 128, 6); lin("vis out", Mv, 1024, 1024, 6)
lin("vis ffn1", Mv, 1024, 1024, 12); lin("vis ffn2", Mv, 1024, 1024, 12)
lin("bi qkv v", Mv, 3072, 1024, 6); lin("bi qkv t", Mt, 3072, 768, 6)
attn("bi t<-v", P, 8, 80, 288, 128, 6); attn("bi v<-t", P, 8, 288, 80, 128, 6)
lin("bi dense1", Mv, 1024, 1024, 6); lin("bi dense2", Mt, 768, 1024, 6)
lin("pool t", P, 1024, 768, 1); lin("pool v", P, 1024, 1024, 1)
lin("lm transform", Mt, 768, 768, 1); lin("lm decoder", Mt, 30522, 768, 1)
lin("img transform", Mv, 1024, 1024, 1); lin("img decoder", Mv, 1601, 1024, 1)
tot_full = tot_q = 0.0; n = 0
rows = []
for name, M, N, K, B, cnt in L:
    tiles = math.ceil(M / 128) * math.ceil(N / 128) * B
    k16 = math.ceil(K / 16)
    cyc_tile = k16 * CYC_PER_K16
    full = tiles * cyc_tile / SMS / CLK * cnt
    quant = math.ceil(tiles / SMS) * cyc_tile / CLK * cnt
    pad = (math.ceil(M / 128) * 128 * math.ceil(N / 128) * 128) / (M * N)
    tot_full += full; tot_q += quant; n += cnt
    rows.append((quant, name, M, N, K, B, cnt, tiles, full, pad))
print(f"{n} launches/step; MMA time at full fill {tot_full*1e3:.2f} ms, with wave quantisation {tot_q*1e3:.2f} ms")
for quant, name, M, N, K, B, cnt, tiles, full, pad in sorted(rows, reverse=True)[:40]:
    print(f"{name:18s} M={M:5d} N={N:5d} K={K:5d} B={B:3d} x{cnt:2d} tiles={tiles:5d} pad={pad:4.2f}  full {full*1e6:7.1f} us  quantised {quant*1e6:7.1f} us")
