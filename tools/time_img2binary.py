import importlib, sys, time
sys.path.insert(0, '/root/repo')
import bench, torch
pkg = importlib.import_module(bench.PKG)
n = 296
eng = pkg.LaneEngine("project", device=0, max_frames=8)
frames = torch.from_numpy(bench.synth_frames(n, 0)).cuda()
def run():
    return eng.grad_binary_dev(frames, eng.gray_dev(frames), s_thresh=(100, 255), sx_thresh=(20, 100), mag_thresh=(50, 255))
run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); run(); e1.record(); torch.cuda.synchronize()
print("img2binary us/frame", 1e3 * e0.elapsed_time(e1) / 2 / n)
