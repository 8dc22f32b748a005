"""decode_probe for an arbitrary tile count (cols x rows): python scripts/decode_probe_n.py <cols> <rows>"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
import libheif_b200 as lb
cols, rows = int(sys.argv[1]), int(sys.argv[2])
tiles = bench.make_tiles(range(cols * rows))
dec = lb.Decoder(host_threads=16)
for _ in range(3):
    dec.decode_grid(tiles, cols=cols, rows=rows)
    out = dec.to_rgb_device(lb.CHROMA_INTERLEAVED_RGB)
    torch.cuda.synchronize()
    st = dec.stats()
print(f"{cols * rows} tiles: entropy {st.entropy_ms:.2f} recon {st.recon_ms:.2f} gpu {st.gpu_ms:.2f} ms")
