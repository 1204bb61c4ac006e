import torch, sys
torch.manual_seed(0)
n = torch.randn(10000, 3, device="cuda")
V = torch.randn(4, 4, device="cuda")
xyz = torch.randn(10000, 3, device="cuda")
outs = []
for i in range(4):
    a = n @ V[:3, :3]
    d = -(n * xyz).sum(-1)
    ld = d - torch.sum(a * V[[3], :3], dim=1)
    outs.append((a.clone(), ld.clone()))
torch.cuda.synchronize()
for i in range(1, 4):
    print(i, "matmul equal:", torch.equal(outs[0][0], outs[i][0]), "max diff", (outs[0][0] - outs[i][0]).abs().max().item(),
          "| dist equal:", torch.equal(outs[0][1], outs[i][1]))
