"""How much of a training step is host (Python + ctypes) time?  Wall time of enqueueing one step
without waiting for the GPU vs the device time of the same step."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ct_image_segmentation_b200 as B
torch.manual_seed(0)
net = B.UNet(3, 1, 10, [16, 32, 64, 128, 256], [2, 2, 2, 2], num_res_units=2).cuda()
loss_fx = B.DiceLoss(include_background=False, to_onehot_y=True, softmax=True)
x = torch.randn(2, 1, 128, 128, 128, device="cuda")
lab = torch.randint(0, 10, (2, 128, 128, 128), device="cuda", dtype=torch.uint8)
def step():
    for p in net.parameters(): p.grad = None
    loss_fx(net(x), lab.unsqueeze(1)).backward()
for _ in range(3): step()
torch.cuda.synchronize()
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0 = time.perf_counter(); e0.record()
    step()
    t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"host enqueue {1e3*(t1-t0):.2f} ms, device {e0.elapsed_time(e1):.2f} ms, wall {1e3*(t2-t0):.2f} ms")
