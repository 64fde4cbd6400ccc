import sys, torch
sys.path.insert(0, '.')
from cassnat_asr_public_b200 import ops
DEV = 'cuda:0'
B, Tin, Fin, C = 64, 2000, 80, 512
x = torch.randn(B, Tin, Fin, device=DEV)
conv1 = torch.nn.Conv2d(1, C, 3, 2, 1).to(DEV); conv2 = torch.nn.Conv2d(C, C, 3, 2, 1).to(DEV)
with torch.no_grad():
    for _ in range(3): y = ops.conv_frontend(x, conv1.weight, conv1.bias, conv2.weight, conv2.bias)
torch.cuda.synchronize()
print("ok", tuple(y.shape))
