// Layout of the three opaque rasterizer buffers (geom / binning / image) -- shared by forward and backward.
//
// HBM layout (all arrays 256-B aligned inside their buffer):
//   geom    (per Gaussian, P rows)  : xy f32x2 | depth f32 | conic+opacity f32x4 | rgb f32x3 | cov3D f32x6 | tiles u32 | tile rect u16x4 |
//                                     clamp mask u32 | depth key u32 | identity idx u32 | depth-sorted key/idx u32 |
//                                     instance offsets (depth order) u32 | scan partials | sort workspace | counters
//   binning (per instance, D rows)  : tile id u32 x2 (ping-pong) | gaussian id u32 x2 | 48-B blend record | sort workspace
//   image   (per pixel / per tile)  : final_T f32 | n_contrib u32 | tile range u32x2 | tile work order u32
// The 48-B record is what the tile kernels stream through shared memory with bulk (TMA) copies:
//   { x, y, conic.x, conic.y | conic.z, opacity, r, g | b, gaussian id (bits), 0, 0 }
#pragma once
#include "common.cuh"
#include "sort_scan.cuh"

namespace mb {

enum Counter { kCntRendered = 0, kCntVisible = 1, kCntOverflow = 2, kCntTileCursor = 3, kCntBwdCursor = 4, kNumCounters = 16 };

struct Record {   // 48 bytes, 16-B aligned
    float4 a;     // x, y, conic.x, conic.y
    float4 b;     // conic.z, opacity, r, g
    float4 c;     // b, id-as-float-bits, 0, 0
};
static_assert(sizeof(Record) == 48, "record must be 48 bytes");

struct GeomState {
    uint32_t *counters;
    float2 *xy;
    float *depth;
    float4 *conic_opacity;
    float *rgb;        // [P*3]
    float *cov3D;      // [P*6]
    uint32_t *tiles_touched;
    ushort4 *rect;     // tile rectangle (x0, y0, x1, y1), exclusive upper bounds
    uint32_t *clamped;
    uint32_t *depth_key, *ident, *sorted_key, *sorted_idx, *offsets, *scan_partials;
    void *sort_ws;
    size_t bytes;

    static GeomState carve(void *p, int64_t P) {
        Carver c(p);
        GeomState g;
        const size_t n = (size_t)(P > 0 ? P : 1);
        g.counters = c.take<uint32_t>(kNumCounters);
        g.xy = c.take<float2>(n);
        g.depth = c.take<float>(n);
        g.conic_opacity = c.take<float4>(n);
        g.rgb = c.take<float>(3 * n);
        g.cov3D = c.take<float>(6 * n);
        g.tiles_touched = c.take<uint32_t>(n);
        g.rect = c.take<ushort4>(n);
        g.clamped = c.take<uint32_t>(n);
        g.depth_key = c.take<uint32_t>(n);
        g.ident = c.take<uint32_t>(n);
        g.sorted_key = c.take<uint32_t>(n);
        g.sorted_idx = c.take<uint32_t>(n);
        g.offsets = c.take<uint32_t>(n);
        g.scan_partials = c.take<uint32_t>((size_t)scan_blocks((int64_t)n) + 1);
        g.sort_ws = c.take<char>(sort_workspace_bytes((int64_t)n));
        g.bytes = c.off;
        return g;
    }
};

struct BinningState {
    uint32_t *tile_a, *tile_b;   // tile id per instance (ping-pong)
    uint32_t *gid_a, *gid_b;     // gaussian id per instance
    Record *records;             // sorted by (tile, depth)
    void *sort_ws;
    size_t bytes;

    static BinningState carve(void *p, int64_t capacity) {
        Carver c(p);
        BinningState b;
        const size_t n = (size_t)(capacity > 0 ? capacity : 1);
        b.tile_a = c.take<uint32_t>(n);
        b.tile_b = c.take<uint32_t>(n);
        b.gid_a = c.take<uint32_t>(n);
        b.gid_b = c.take<uint32_t>(n);
        b.records = c.take<Record>(n + 1);
        b.sort_ws = c.take<char>(sort_workspace_bytes((int64_t)n));
        b.bytes = c.off;
        return b;
    }
};

struct ImageState {
    float *final_T;        // [H*W]
    uint32_t *n_contrib;   // [H*W]
    uint2 *ranges;         // [tiles] (start, end) into records
    uint32_t *tile_order;  // [tiles] tiles sorted by descending length (work queue)
    uint32_t *tile_len;    // [tiles]
    size_t bytes;

    static ImageState carve(void *p, int W, int H) {
        Carver c(p);
        ImageState s;
        const size_t px = (size_t)W * H;
        const size_t tiles = (size_t)((W + kTile - 1) / kTile) * ((H + kTile - 1) / kTile);
        s.final_T = c.take<float>(px ? px : 1);
        s.n_contrib = c.take<uint32_t>(px ? px : 1);
        s.ranges = c.take<uint2>(tiles ? tiles : 1);
        s.tile_order = c.take<uint32_t>(tiles ? tiles : 1);
        s.tile_len = c.take<uint32_t>(tiles ? tiles : 1);
        s.bytes = c.off;
        return s;
    }
};

// validated view of mb_raster_inputs
struct RasterDims {
    int P, W, H, gx, gy, tiles;
    float focx, focy;
};

inline RasterDims raster_dims(const mb_raster_inputs *in) {
    RasterDims d;
    d.P = in->num_points;
    d.W = in->image_width;
    d.H = in->image_height;
    d.gx = (d.W + kTile - 1) / kTile;
    d.gy = (d.H + kTile - 1) / kTile;
    d.tiles = d.gx * d.gy;
    d.focx = d.W / (2.0f * in->tanfovx);
    d.focy = d.H / (2.0f * in->tanfovy);
    return d;
}

int validate_raster_inputs(const mb_raster_inputs *in, const char *who);

}  // namespace mb
