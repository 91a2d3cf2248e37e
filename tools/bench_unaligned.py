import sys, torch, statistics
sys.path.insert(0,'/root/repo')
import rustyhgi_b200 as hgi
def t(fn,n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); b.synchronize(); return a.elapsed_time(b)/n
for w in (1920, 1916, 1918, 1919):
    frames = torch.randint(0,256,(256,1080,w),dtype=torch.uint8,device='cuda')
    for name,path in (("swar",hgi.PATH_TILE),("generic",hgi.PATH_TILE_GENERIC)):
        ctx=hgi.Context(0,path)
        enc=hgi.Encoder(hgi.Crossed,hgi.Linear(hgi.QuantizationLevel.Medium),4,ctx=ctx); dec=hgi.Decoder(hgi.Crossed,ctx=ctx)
        g=torch.empty_like(frames); o=torch.empty_like(frames)
        te=t(lambda: enc.encode_device(frames,grids_out=g)); td=t(lambda: dec.decode_device(4,g,images_out=o))
        print(w,name,f"enc {te:.3f} ms dec {td:.3f} ms  -> {frames.numel()/ (te+td)/1e3:.0f} Mpx/s")
        ctx.close()
