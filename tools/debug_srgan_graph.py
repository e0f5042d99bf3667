"""Debug helper: capture parts of the SRGAN adversarial step in a CUDA graph and replay them (one variant per process)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pytorch-super-resolution-model-collection_b200"))

VARIANTS = ["b_all3_pool", "b_all3_nopool", "b_pinned_pool"]


def run(variant):
    import torch
    import srb200
    from srb200 import host, nn_ops
    from srb200 import functional as F
    dev = torch.device("cuda", 0)
    benchlike = variant.startswith("b_")
    opts = variant
    if benchlike and opts != "b_nomath":
        srb200.set_math("auto")
    torch.manual_seed(0)
    G = srb200.models.SRGANGenerator(3, 64, 16)
    host.init_model("srgan", G)
    if benchlike and opts != "b_seeds":
        torch.manual_seed(1)
    D = srb200.models.SRGANDiscriminator(3, 64, 128)
    host.init_model("srgan", D)
    if benchlike and opts != "b_seeds":
        torch.manual_seed(2)
    FE = srb200.models.FeatureExtractor()
    if benchlike and opts != "b_nokaiming":
        for m in FE.modules():
            if isinstance(m, torch.nn.Conv2d):
                torch.nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                torch.nn.init.constant_(m.bias, 0)
    for m in (G, D, FE):
        m.to(dev).train()
    go, do = host.make_srgan_optimizers(G, D, lr=1e-5, capturable=True)
    bg = bd = None
    three = variant.startswith("full3") or "3_" in variant or "pinned" in variant
    use_pool = variant.endswith("_pool")
    if "pinned" in opts:
        keep = [torch.randint(0, 256, (16, 128, 128, 3), dtype=torch.uint8).pin_memory() for _ in range(6)]
    if three:
        variant = "full_bucket" if "bucket" in variant else "full"
    if variant == "full_bucket":
        bg, bd = srb200.GradBucket(G, world_size=1), srb200.GradBucket(D, world_size=1)
    lr_img = torch.rand(16, 3, 32, 32, device=dev)
    hr_img = torch.rand(16, 3, 128, 128, device=dev)
    if benchlike and opts != "b_randinput":
        gen = torch.Generator().manual_seed(1)
        lr_img = srb200.image_to_tensor(torch.randint(0, 256, (16, 32, 32, 3), generator=gen, dtype=torch.uint8).to(dev))
        hr_img = srb200.image_to_tensor(torch.randint(0, 256, (16, 128, 128, 3), generator=gen, dtype=torch.uint8).to(dev))
    nwarm = 3 if (benchlike and opts != "b_warm2") else 2
    if benchlike:
        variant = "full"
    ones = torch.ones(16, device=dev)

    def step():
        if variant in ("full", "full_bucket"):
            dl, gl = host.srgan_step(G, D, FE, go, do, lr_img, hr_img, bucket_g=bg, bucket_d=bd)
            return dl + gl
        x_, y_ = host.norm_vgg(hr_img), host.norm_vgg(lr_img)
        if variant == "d_fwd":
            with torch.no_grad():
                return nn_ops.bce_loss(D(x_), ones)
        if variant == "g_fwd":
            with torch.no_grad():
                return F.mse_loss(G(y_), x_)
        if variant in ("d_only", "d_fwd_bwd"):
            do.zero_grad()
            loss = nn_ops.bce_loss(D(x_), ones)
            loss.backward()
            if variant == "d_only":
                do.step()
            return loss.detach()
        if variant == "g_only":
            go.zero_grad()
            loss = F.mse_loss(G(y_), x_)
            loss.backward()
            go.step()
            return loss.detach()
        if variant == "fe_only":
            a = FE(x_)
            loss = F.mse_loss(a, torch.zeros_like(a))
            loss.backward()
            return loss.detach()
        if variant == "full_noopt":
            do.zero_grad(); go.zero_grad()
            recon = G(y_)
            loss = nn_ops.bce_loss(D(recon), ones) + F.mse_loss(recon, x_)
            loss.backward()
            return loss.detach()
        raise SystemExit("unknown variant")

    for _ in range(nwarm):
        step()
    torch.cuda.synchronize()
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    if three:
        pool = torch.cuda.graph_pool_handle() if use_pool else None
        gs, outs = [], []
        with torch.cuda.stream(side):
            for i in range(3):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool, stream=side):
                    out = step().clone()
                gs.append(g); outs.append(out)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize()
        print(variant, "captured x3", flush=True)
        for k in range(6):
            gs[k % 3].replay()
            torch.cuda.synchronize()
            print(variant, "replay", k, float(outs[k % 3]), flush=True)
        print(variant, "replayed OK", flush=True)
        return
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            out = step().clone()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    print(variant, "captured", flush=True)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    print(variant, "replayed OK", float(out), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(sys.argv[1])
    else:
        for v in VARIANTS:
            r = subprocess.run([sys.executable, "-X", "faulthandler", __file__, v], capture_output=True, text=True, timeout=300)
            tail = (r.stdout.strip().splitlines() or [""])[-1]
            print("%-12s rc=%d  %s" % (v, r.returncode, tail), flush=True)
            if r.returncode != 0:
                print("    " + "\n    ".join(r.stderr.strip().splitlines()[-12:]), flush=True)
