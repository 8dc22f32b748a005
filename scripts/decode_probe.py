"""Small decode workload for ncu captures: side x side tiles of 1024x1024 through the device front-end."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
import libheif_b200 as lb
side = int(sys.argv[1]) if len(sys.argv) > 1 else 4
fe = sys.argv[2] if len(sys.argv) > 2 else "device"
tiles = bench.make_tiles(range(side * side))
dec = lb.Decoder(host_threads=16)
dec.set_front_end(fe == "device")
for _ in range(3):
    dec.decode_grid(tiles, cols=side, rows=side)
    out = dec.to_rgb_device(lb.CHROMA_INTERLEAVED_RGB)
    torch.cuda.synchronize()
    st = dec.stats()
    print(f"entropy {st.entropy_ms:.3f} recon {st.recon_ms:.3f} deblock {st.deblock_ms:.3f} sao {st.sao_ms:.3f} parse {st.parse_ms:.2f} ms")
