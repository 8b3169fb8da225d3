import torch
A = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
B = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
for _ in range(3):
    C = torch.matmul(A, B)
torch.cuda.synchronize()
print("ok", float(C[0, 0]))
