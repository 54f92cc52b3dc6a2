"""Seeded synthetic cohorts generated on the GPU with torch (bench / large-test plumbing only).

Same distributions as bedio.synth_packed: MAF_k ~ U(maf_lo, maf_hi), dosage ~
Binomial(2, MAF_k), i.i.d. missingness.  Returns the packed SNP-major .bed rows as a CUDA uint8
tensor plus the per-SNP sample MAF (for the `--extract`-list MAF filter)."""
import torch


def synth_packed_cuda(n, m, seed, maf_lo=0.05, maf_hi=0.5, missing=0.0, device="cuda", rows_per_block=2048):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    b = (n + 3) // 4
    out = torch.empty((m, b), dtype=torch.uint8, device=device)
    mafs = torch.empty(m, dtype=torch.float64, device=device)
    table = torch.tensor([0b00, 0b10, 0b11, 0b01], dtype=torch.uint8, device=device)  # dosage 0/1/2, 3 = missing
    shifts = torch.tensor([1, 4, 16, 64], dtype=torch.int32, device=device)
    for r0 in range(0, m, rows_per_block):
        r1 = min(m, r0 + rows_per_block)
        rows = r1 - r0
        maf = torch.rand((rows, 1), generator=g, device=device) * (maf_hi - maf_lo) + maf_lo
        dos = (torch.rand((rows, n), generator=g, device=device) < maf).to(torch.int64)
        dos += (torch.rand((rows, n), generator=g, device=device) < maf).to(torch.int64)
        if missing > 0:
            dos[torch.rand((rows, n), generator=g, device=device) < missing] = 3
        # keep every SNP polymorphic (the reference aborts on standardDev == 0)
        present = dos != 3
        nn = present.sum(dim=1).clamp(min=1)
        tot = torch.where(present, dos, torch.zeros_like(dos)).sum(dim=1)
        mono = (tot == 0) | (tot == 2 * nn)
        if n >= 2 and bool(mono.any()):
            dos[mono, 0] = 0
            dos[mono, 1] = 2
            present = dos != 3
            nn = present.sum(dim=1).clamp(min=1)
            tot = torch.where(present, dos, torch.zeros_like(dos)).sum(dim=1)
        p = tot.to(torch.float64) / (2.0 * nn.to(torch.float64))
        mafs[r0:r1] = torch.minimum(p, 1.0 - p)
        pat = table[dos].to(torch.int32)
        if n % 4:
            pad = torch.zeros((rows, 4 - n % 4), dtype=torch.int32, device=device)
            pat = torch.cat([pat, pad], dim=1)
        out[r0:r1] = (pat.view(rows, b, 4) * shifts).sum(dim=2).to(torch.uint8)
    return out, mafs
