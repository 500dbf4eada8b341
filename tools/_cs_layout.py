"""Scratch-buffer layout of the tensor-core Chamfer path (mirror of cs_layout in csrc/chamfer_sweep.cu),
for the measurement tools that peek into the buffer after a call.  Not part of the product."""
R2_SLOTS = 64
up = lambda x: (x + 255) & ~255


def layout(B, N, M):
    o = up(4 * B * (R2_SLOTS + 1)); L = {"r2part": 0, "taubits": 4 * B * R2_SLOTS}
    for s, n in enumerate((N, M)):
        rows = B * ((n + 127) // 128) * 128
        L["aform%d" % s] = o; o += up(64 * rows)
        L["bform%d" % s] = o; o += up(64 * rows)
        L["norm%d" % s] = o; o += up(4 * rows)
        L["key%d" % s] = o; o += up(8 * B * n)
        L["sec%d" % s] = o; o += up(4 * B * n)
        L["mask%d" % s] = o; o += up(8 * B * n)
    L["total"] = o
    return L


def tau(ws, B):
    """Per cloud pair: the ambiguity margin TAU the sweep used (float32 tensor of B)."""
    import torch
    off = 4 * B * R2_SLOTS
    return ws[off:off + 4 * B].view(torch.float32)


def ambiguous(ws, B, N, M):
    """Counts of points (per direction) whose runner-up granule lies within TAU of the best value."""
    import torch
    L = layout(B, N, M); t = tau(ws, B).view(B, 1); out = []
    for s, n in enumerate((N, M)):
        key = ws[L["key%d" % s]:L["key%d" % s] + 8 * B * n].view(torch.int64).view(B, n)
        best = (key >> 32).to(torch.int32).view(torch.float32)
        sec = ws[L["sec%d" % s]:L["sec%d" % s] + 4 * B * n].view(torch.float32).view(B, n)
        out.append(int((sec <= best + t).sum()))
    return out
