"""Developer measurement (GPU): how many (slice, column-tile, k-block) boxes of the sliced factor R^T (forward) / R (backward) are
all-zero, i.e. how much TMA + UMMA work a zero-skipping contraction could save."""
import sys
import torch
sys.path.insert(0, ".")
from botorch_b200.benchmarks import configs
dev = torch.device("cuda:0")
for cfg in sys.argv[1:] or ["C2", "C3"]:
    data = configs.make_problem(configs.CONFIGS[cfg]); model = configs.build_model(data, dev)
    st = model.prediction_strategy()
    for name, S, bn in (("Rt (forward, BN=80)", st.Rt_slices, 80), ("R (backward, BN=96)", st.R_slices, 96)):
        G, N, K = S.shape
        nt, kb = (N + bn - 1) // bn, K // 64
        pad = nt * bn - N
        Sp = torch.nn.functional.pad(S, (0, 0, 0, pad)) if pad else S
        nz = (Sp.view(G, nt, bn, kb, 64) != 0).any(dim=4).any(dim=2)       # [G][nt][kb]
        # triangular structure: only k-blocks inside the non-zero k-range of the tile are visited at all
        visited = nz.any(dim=0)
        lead = torch.zeros(nt, kb, dtype=torch.long, device=dev)          # number of leading all-zero slices per box
        alive = torch.ones(nt, kb, dtype=torch.bool, device=dev)
        for g in range(G):
            alive &= ~nz[g]
            lead += alive.long()
        v = visited.sum().item()
        fr = [float(((~nz[g]) & visited).sum().item()) / v for g in range(G)]
        hist = [int(((lead == z) & visited).sum().item()) for z in range(G + 1)]
        # products saved if slices 0..z-1 of a box are skipped: sum over p of min(z, G - p) of G (G + 1) / 2
        total = G * (G + 1) // 2
        saved = sum(h * sum(min(z, G - p) for p in range(G)) for z, h in enumerate(hist)) / (v * total)
        print(f"{cfg} {name}: G={G}, visited boxes {v} of {nt * kb}; zero fraction per slice {['%.2f' % f for f in fr]}; "
              f"leading-zero-slice histogram {hist}; slice products saved by skipping leading zero slices: {100 * saved:.1f} %")
