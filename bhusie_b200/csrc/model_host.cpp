// model_host.cpp — host-side scene preparation behind bh_model_* (include/bh_abi.h).
//
// Replaces, for the ray pass's `models` storage buffer (ray.wgsl:9,53-65):
//   load_model            src/renderer/model.rs:7-87        OBJ -> points/normals/triangles
//   Model::build_bvh      src/renderer/triangle.rs:143-157  root + recursion entry
//   Model::update_bounds  src/renderer/triangle.rs:159-194
//   Model::subdivide      src/renderer/triangle.rs:196-259  midpoint split, in-place partition
//   ModelUniform          src/renderer/triangle.rs:268-285  byte layout written here verbatim
//
// BVH hit indices must be bit-exact with the reference, so node numbering and bvh_lookup order
// are reproduced exactly: children are allocated as a consecutive pair at the moment their parent
// splits, and the left subtree is completely built before the right one (pre-order).  The build
// here is iterative (explicit stack) instead of the reference's recursion, which needs a 1 GiB
// thread stack (src/main.rs:1-5).
#include <cerrno>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include <zlib.h>

#include "../../include/bh_abi.h"

namespace bh {
void set_error(const char *fmt, ...);   // bh_abi.cu
}

namespace {

constexpr size_t kMuPoints = 48, kMuNormals = 8388656, kMuTriangles = 16777264, kMuNodes = 29360176, kMuLookup = 46137392;

struct NodeU { float mn[3]; int32_t left_child; float mx[3]; int32_t obj_count; };
struct TriU { int32_t p1, p2, p3, n1, n2, n3; };
static_assert(sizeof(NodeU) == 32 && sizeof(TriU) == 24, "ModelUniform element sizes");

struct ModelView {
    float *points; float *normals; TriU *tris; NodeU *nodes; int32_t *lookup;
    explicit ModelView(void *blob)
    {
        auto *b = static_cast<unsigned char *>(blob);
        points = reinterpret_cast<float *>(b + kMuPoints);
        normals = reinterpret_cast<float *>(b + kMuNormals);
        tris = reinterpret_cast<TriU *>(b + kMuTriangles);
        nodes = reinterpret_cast<NodeU *>(b + kMuNodes);
        lookup = reinterpret_cast<int32_t *>(b + kMuLookup);
    }
};

void fit_bounds(ModelView &m, NodeU &node)
{
    for (int a = 0; a < 3; ++a) { node.mn[a] = 3.40282347e+38f; node.mx[a] = -3.40282347e+38f; }
    for (int32_t k = 0; k < node.obj_count; ++k) {
        const TriU &t = m.tris[m.lookup[node.left_child + k]];
        for (int32_t pi : { t.p1, t.p2, t.p3 }) {
            const float *p = m.points + 4 * static_cast<size_t>(pi);
            for (int a = 0; a < 3; ++a) {
                node.mn[a] = fminf(node.mn[a], p[a]);
                node.mx[a] = fmaxf(node.mx[a], p[a]);
            }
        }
    }
}

int build_bvh(void *blob, int32_t triangle_count, bh_model_info *info)
{
    if (!blob || triangle_count < 0 || triangle_count > BH_MAX_MODEL_VERTICES) {
        bh::set_error("bh_model_build_bvh: bad arguments (triangle_count=%d)", triangle_count);
        return BH_ERR_INVALID;
    }
    ModelView m(blob);
    for (int32_t i = 0; i < triangle_count; ++i) m.lookup[i] = i;
    size_t used = 1;
    m.nodes[0].left_child = 0;
    m.nodes[0].obj_count = triangle_count;
    fit_bounds(m, m.nodes[0]);

    struct Pending { int32_t node; int32_t depth; };
    std::vector<Pending> todo;
    todo.push_back({ 0, 0 });
    int32_t max_depth = 0;
    while (!todo.empty()) {
        const Pending cur = todo.back();
        todo.pop_back();
        if (cur.depth > max_depth) max_depth = cur.depth;
        NodeU &node = m.nodes[cur.node];
        if (node.obj_count <= 2) continue;
        float extent[3];
        for (int a = 0; a < 3; ++a) extent[a] = node.mx[a] - node.mn[a];
        int axis = 0;
        if (extent[1] > extent[axis]) axis = 1;
        if (extent[2] > extent[axis]) axis = 2;
        const float split = node.mn[axis] + extent[axis] / 2.0f;
        int32_t lo = node.left_child;
        int32_t hi = lo + node.obj_count - 1;
        while (lo <= hi) {
            const TriU &t = m.tris[m.lookup[lo]];
            const float centroid = (m.points[4 * static_cast<size_t>(t.p1) + axis] + m.points[4 * static_cast<size_t>(t.p2) + axis] +
                                    m.points[4 * static_cast<size_t>(t.p3) + axis]) / 3.0f;
            if (centroid < split) {
                ++lo;
            } else {
                std::swap(m.lookup[lo], m.lookup[hi]);
                --hi;
            }
        }
        const int32_t left_count = lo - node.left_child;
        if (left_count == 0 || left_count == node.obj_count) continue;
        if (used + 2 > static_cast<size_t>(BH_MAX_MODEL_VERTICES)) {
            bh::set_error("bh_model_build_bvh: node array exhausted");
            return BH_ERR_TOOBIG;
        }
        const int32_t left = static_cast<int32_t>(used++), right = static_cast<int32_t>(used++);
        m.nodes[left].left_child = node.left_child;
        m.nodes[left].obj_count = left_count;
        m.nodes[right].left_child = lo;
        m.nodes[right].obj_count = node.obj_count - left_count;
        node.left_child = left;
        node.obj_count = 0;
        fit_bounds(m, m.nodes[left]);
        fit_bounds(m, m.nodes[right]);
        todo.push_back({ right, cur.depth + 1 });   // popped after the whole left subtree
        todo.push_back({ left, cur.depth + 1 });
    }
    if (info) {
        info->triangle_count = triangle_count;
        info->nodes_used = static_cast<int32_t>(used);
        info->max_depth = max_depth;
        int32_t leaves = 0, biggest = 0;
        for (size_t i = 0; i < used; ++i)
            if (m.nodes[i].obj_count > 0) { ++leaves; if (m.nodes[i].obj_count > biggest) biggest = m.nodes[i].obj_count; }
        info->leaf_count = leaves;
        info->max_leaf_size = biggest;
    }
    return BH_OK;
}

void write_header(void *blob, const float position[3], int32_t visible, int32_t point_count, int32_t triangle_count)
{
    auto *b = static_cast<unsigned char *>(blob);
    std::memcpy(b + 0, position, 12);
    std::memcpy(b + 12, &visible, 4);
    std::memcpy(b + 32, &point_count, 4);
    // normal_count @36 is never written by ModelUniform::update (triangle.rs:309-324, Q17)
    std::memcpy(b + 40, &triangle_count, 4);
}

// one "model" in tobj's sense: first-use re-indexing of positions and normals
struct ObjectState {
    std::unordered_map<int32_t, int32_t> pos_map, nrm_map;
    bool has_faces = false;
    void reset() { pos_map.clear(); nrm_map.clear(); has_faces = false; }
};

const char *skip_ws(const char *p) { while (*p == ' ' || *p == '\t') ++p; return p; }

}  // namespace

// Walks the BVH of a ModelUniform blob from its root the way trace_ray_model (ray.wgsl:287-363) can reach it and checks
// every index the kernel would follow: children inside the node array and numbered after their parent (which makes the
// traversal finite), leaf ranges inside bvh_lookup, lookup entries inside the triangle array, point / normal indices inside
// their arrays.  The reference's WGSL clamps out-of-range indices (naga Restrict); this library refuses the blob instead.
extern "C" int bh_model_validate(const void *model_uniform)
{
    if (!model_uniform) { bh::set_error("bh_model_validate: null argument"); return BH_ERR_INVALID; }
    ModelView m(const_cast<void *>(model_uniform));
    constexpr int32_t kMax = BH_MAX_MODEL_VERTICES;
    // an empty model (Model::new + build_bvh over 0 triangles, triangle.rs:143-157) has root = {left_child 0, obj_count 0}: the
    // shader reads it as an inner node whose children are nodes 0 and 1; node 0 then carries the builder's empty bounds
    // (min = f32::MAX, max = f32::MIN, triangle.rs:160-161), which no ray hits, so the traversal ends at once — kept as is
    if (m.nodes[0].obj_count == 0 && m.nodes[0].left_child == 0 && m.nodes[0].mn[0] > m.nodes[0].mx[0]) return BH_OK;
    std::vector<int32_t> todo;
    todo.push_back(0);
    size_t visited = 0;
    while (!todo.empty()) {
        const int32_t ni = todo.back();
        todo.pop_back();
        if (++visited > static_cast<size_t>(kMax)) { bh::set_error("bh_model_validate: BVH has more reachable nodes than the node array"); return BH_ERR_INVALID; }
        const NodeU &node = m.nodes[ni];
        if (node.obj_count == 0) {
            const int32_t c = node.left_child;
            if (c <= ni || c < 0 || c > kMax - 2) { bh::set_error("bh_model_validate: node %d has children %d,%d (must follow their parent inside the array)", ni, c, c + 1); return BH_ERR_INVALID; }
            todo.push_back(c + 1);
            todo.push_back(c);
        } else {
            if (node.obj_count < 0 || node.left_child < 0 || node.left_child > kMax - node.obj_count) {
                bh::set_error("bh_model_validate: leaf %d covers bvh_lookup[%d..+%d)", ni, node.left_child, node.obj_count);
                return BH_ERR_INVALID;
            }
            for (int32_t k = 0; k < node.obj_count; ++k) {
                const int32_t ti = m.lookup[node.left_child + k];
                if (ti < 0 || ti >= kMax) { bh::set_error("bh_model_validate: bvh_lookup[%d] = %d", node.left_child + k, ti); return BH_ERR_INVALID; }
                const TriU &t = m.tris[ti];
                for (int32_t idx : { t.p1, t.p2, t.p3, t.n1, t.n2, t.n3 })
                    if (idx < 0 || idx >= kMax) { bh::set_error("bh_model_validate: triangle %d has index %d", ti, idx); return BH_ERR_INVALID; }
            }
        }
    }
    return BH_OK;
}

extern "C" int bh_model_build_bvh(void *model_uniform, int32_t triangle_count, bh_model_info *info)
{
    if (info) std::memset(info, 0, sizeof *info);
    return build_bvh(model_uniform, triangle_count, info);
}

extern "C" int bh_model_from_arrays(const float *points, int32_t n_points, const float *normals, int32_t n_normals,
                                    const int32_t *tris, int32_t n_tris, const float position[3], int32_t visible,
                                    void *model_uniform, bh_model_info *info)
{
    if (info) std::memset(info, 0, sizeof *info);
    if (!points || !normals || !tris || !position || !model_uniform || n_points < 0 || n_normals < 0 || n_tris < 0) {
        bh::set_error("bh_model_from_arrays: null or negative argument");
        return BH_ERR_INVALID;
    }
    if (n_points > BH_MAX_MODEL_VERTICES || n_normals > BH_MAX_MODEL_VERTICES || n_tris > BH_MAX_MODEL_VERTICES) {
        bh::set_error("bh_model_from_arrays: mesh exceeds MAX_MODEL_VERTICES=%d", BH_MAX_MODEL_VERTICES);
        return BH_ERR_TOOBIG;
    }
    std::memset(model_uniform, 0, BH_MODEL_UNIFORM_SIZE);
    ModelView m(model_uniform);
    for (int32_t i = 0; i < n_points; ++i) {
        m.points[4 * i + 0] = points[3 * i + 0]; m.points[4 * i + 1] = points[3 * i + 1]; m.points[4 * i + 2] = points[3 * i + 2];
    }
    for (int32_t i = 0; i < n_normals; ++i) {
        m.normals[4 * i + 0] = normals[3 * i + 0]; m.normals[4 * i + 1] = normals[3 * i + 1]; m.normals[4 * i + 2] = normals[3 * i + 2];
    }
    for (int32_t i = 0; i < n_tris; ++i) {
        const int32_t *t = tris + 6 * static_cast<size_t>(i);
        for (int k = 0; k < 3; ++k) {
            if (t[k] < 0 || t[k] >= n_points || t[3 + k] < 0 || t[3 + k] >= n_normals) {
                bh::set_error("bh_model_from_arrays: triangle %d has an out-of-range index", i);
                return BH_ERR_INVALID;
            }
        }
        m.tris[i] = TriU{ t[0], t[1], t[2], t[3], t[4], t[5] };
    }
    write_header(model_uniform, position, visible, n_points, n_tris);
    const int rc = build_bvh(model_uniform, n_tris, info);
    if (rc == BH_OK && info) { info->point_count = n_points; info->normal_count = n_normals; }
    return rc;
}

extern "C" int bh_model_load_obj(const char *path, void *model_uniform, bh_model_info *info)
{
    if (info) std::memset(info, 0, sizeof *info);
    if (!path || !model_uniform) { bh::set_error("bh_model_load_obj: null argument"); return BH_ERR_INVALID; }
    FILE *f = std::fopen(path, "rb");
    if (!f) { bh::set_error("bh_model_load_obj: cannot open %s", path); return BH_ERR_NOENT; }
    std::string text;
    {
        std::fseek(f, 0, SEEK_END);
        const long sz = std::ftell(f);
        std::fseek(f, 0, SEEK_SET);
        text.resize(sz > 0 ? static_cast<size_t>(sz) : 0);
        const size_t got = text.empty() ? 0 : std::fread(&text[0], 1, text.size(), f);
        std::fclose(f);
        if (got != text.size()) { bh::set_error("bh_model_load_obj: short read on %s", path); return BH_ERR_CUDA; }
    }
    std::memset(model_uniform, 0, BH_MODEL_UNIFORM_SIZE);
    ModelView m(model_uniform);

    std::vector<float> file_pos, file_nrm;          // pools in file order
    int32_t point_count = 0, normal_count = 0, triangle_count = 0;
    int32_t mesh_offset = 0, normal_offset = 0;     // model.rs:22-23 (Q17: mesh_offset is a TRIANGLE count)
    ObjectState obj;

    const char *p = text.c_str();
    const char *end = p + text.size();
    while (p < end) {
        const char *line = p;
        while (p < end && *p != '\n') ++p;
        const char *next = p < end ? p + 1 : p;
        const char *q = skip_ws(line);
        if (q[0] == 'v' && (q[1] == ' ' || q[1] == '\t')) {
            char *e = nullptr;
            for (int k = 0; k < 3; ++k) { file_pos.push_back(std::strtof(k ? e : q + 2, &e)); }
        } else if (q[0] == 'v' && q[1] == 'n' && (q[2] == ' ' || q[2] == '\t')) {
            char *e = nullptr;
            for (int k = 0; k < 3; ++k) { file_nrm.push_back(std::strtof(k ? e : q + 3, &e)); }
        } else if ((q[0] == 'o' || q[0] == 'g') && (q[1] == ' ' || q[1] == '\t' || q[1] == '\r' || q[1] == '\n' || q[1] == 0)) {
            if (obj.has_faces) {                      // tobj starts a new model; bhusie offsets it (Q17)
                mesh_offset = triangle_count;
                normal_offset = normal_count;
                obj.reset();
            }
        } else if (q[0] == 'f' && (q[1] == ' ' || q[1] == '\t')) {
            int32_t vi[3], ni[3];
            bool has_n = false;
            const char *c = q + 2;
            int corners = 0;
            const int32_t npos = static_cast<int32_t>(file_pos.size() / 3), nnrm = static_cast<int32_t>(file_nrm.size() / 3);
            while (true) {
                c = skip_ws(c);
                if (*c == '\n' || *c == '\r' || *c == 0) break;
                if (corners == 3) { bh::set_error("bh_model_load_obj: only triangles are supported (%s)", path); return BH_ERR_INVALID; }
                char *e = nullptr;
                const long a = std::strtol(c, &e, 10);
                if (e == c) { bh::set_error("bh_model_load_obj: malformed face in %s", path); return BH_ERR_INVALID; }
                long n = 0; bool hn = false;
                c = e;
                if (*c == '/') {
                    ++c;
                    if (*c != '/') { std::strtol(c, &e, 10); c = e; }
                    if (*c == '/') { ++c; n = std::strtol(c, &e, 10); hn = e != c; c = e; }
                }
                vi[corners] = static_cast<int32_t>(a < 0 ? npos + a : a - 1);
                ni[corners] = hn ? static_cast<int32_t>(n < 0 ? nnrm + n : n - 1) : -1;
                has_n = has_n || hn;
                ++corners;
            }
            if (corners != 3) { bh::set_error("bh_model_load_obj: face with %d corners in %s", corners, path); return BH_ERR_INVALID; }
            if (triangle_count >= BH_MAX_MODEL_VERTICES) { bh::set_error("bh_model_load_obj: too many triangles"); return BH_ERR_TOOBIG; }
            obj.has_faces = true;
            int32_t lp[3], ln[3] = { 0, 0, 0 };
            for (int k = 0; k < 3; ++k) {
                if (vi[k] < 0 || vi[k] >= npos) { bh::set_error("bh_model_load_obj: vertex index out of range"); return BH_ERR_INVALID; }
                auto it = obj.pos_map.find(vi[k]);
                if (it == obj.pos_map.end()) {
                    if (point_count >= BH_MAX_MODEL_VERTICES) { bh::set_error("bh_model_load_obj: too many points"); return BH_ERR_TOOBIG; }
                    const int32_t local = static_cast<int32_t>(obj.pos_map.size());
                    obj.pos_map.emplace(vi[k], local);
                    float *dst = m.points + 4 * static_cast<size_t>(point_count++);
                    dst[0] = file_pos[3 * static_cast<size_t>(vi[k]) + 0] * 0.5f;      // model.rs:36-42
                    dst[1] = file_pos[3 * static_cast<size_t>(vi[k]) + 1] * -0.5f;
                    dst[2] = file_pos[3 * static_cast<size_t>(vi[k]) + 2] * 0.5f;
                    lp[k] = local;
                } else {
                    lp[k] = it->second;
                }
                if (has_n) {
                    if (ni[k] < 0 || ni[k] >= nnrm) { bh::set_error("bh_model_load_obj: normal index out of range"); return BH_ERR_INVALID; }
                    auto jt = obj.nrm_map.find(ni[k]);
                    if (jt == obj.nrm_map.end()) {
                        if (normal_count >= BH_MAX_MODEL_VERTICES) { bh::set_error("bh_model_load_obj: too many normals"); return BH_ERR_TOOBIG; }
                        const int32_t local = static_cast<int32_t>(obj.nrm_map.size());
                        obj.nrm_map.emplace(ni[k], local);
                        float *dst = m.normals + 4 * static_cast<size_t>(normal_count++);
                        dst[0] = file_nrm[3 * static_cast<size_t>(ni[k]) + 0];
                        dst[1] = file_nrm[3 * static_cast<size_t>(ni[k]) + 1];
                        dst[2] = file_nrm[3 * static_cast<size_t>(ni[k]) + 2];
                        ln[k] = local;
                    } else {
                        ln[k] = jt->second;
                    }
                }
            }
            if (!has_n) {
                // model.rs:56-68: per-face normal from model.points[<un-offset index>], appended at normal_count
                if (normal_count >= BH_MAX_MODEL_VERTICES) { bh::set_error("bh_model_load_obj: too many normals"); return BH_ERR_TOOBIG; }
                const float *a = m.points + 4 * static_cast<size_t>(lp[0]);
                const float *b = m.points + 4 * static_cast<size_t>(lp[1]);
                const float *cc = m.points + 4 * static_cast<size_t>(lp[2]);
                const float e1[3] = { b[0] - a[0], b[1] - a[1], b[2] - a[2] }, e2[3] = { cc[0] - a[0], cc[1] - a[1], cc[2] - a[2] };
                const float cr[3] = { e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0] };
                const float inv = 1.0f / sqrtf(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);   // cgmath normalize = v * (1/|v|)
                float *dst = m.normals + 4 * static_cast<size_t>(normal_count);
                dst[0] = cr[0] * inv; dst[1] = cr[1] * inv; dst[2] = cr[2] * inv;
                ln[0] = ln[1] = ln[2] = normal_count++;
            }
            // Q17: mesh_offset is a TRIANGLE count added to POINT indices; on multi-object files the sum can leave the
            // 524288-slot arrays, where the reference shader would clamp (naga Restrict) and this kernel would read out of
            // bounds — refuse instead
            for (int k = 0; k < 3; ++k) {
                if (lp[k] + mesh_offset >= BH_MAX_MODEL_VERTICES || ln[k] + normal_offset >= BH_MAX_MODEL_VERTICES) {
                    bh::set_error("bh_model_load_obj: offset index beyond MAX_MODEL_VERTICES (multi-object OBJ, model.rs:22,71-73)");
                    return BH_ERR_TOOBIG;
                }
            }
            m.tris[triangle_count++] = TriU{ lp[0] + mesh_offset, lp[1] + mesh_offset, lp[2] + mesh_offset,
                                             ln[0] + normal_offset, ln[1] + normal_offset, ln[2] + normal_offset };
        }
        p = next;
    }
    const float position[3] = { -10.0f, 0.0f, 30.0f };    // Model::new, triangle.rs:100
    write_header(model_uniform, position, 1, point_count, triangle_count);
    const int rc = build_bvh(model_uniform, triangle_count, info);
    if (rc == BH_OK && info) { info->point_count = point_count; info->normal_count = normal_count; }
    return rc;
}

// ---------------------------------------------------------------------------------------------------------------------
// Frame dump: the "Save Image" path of Renderer::render (src/renderer/mod.rs:460-486) — the post-FXAA
// RGBA8 frame written as an 8-bit RGBA PNG with alpha forced to 255 (mod.rs:479).  The reference's 256-byte row-pitch
// staging buffer (mod.rs:491-526) is a wgpu copy constraint and has no counterpart here.  zlib is the only dependency.
// ---------------------------------------------------------------------------------------------------------------------
namespace {
void png_chunk(FILE *f, const char type[4], const unsigned char *data, size_t n)
{
    unsigned char len[4] = { (unsigned char)(n >> 24), (unsigned char)(n >> 16), (unsigned char)(n >> 8), (unsigned char)n };
    std::fwrite(len, 1, 4, f);
    std::fwrite(type, 1, 4, f);
    if (n) std::fwrite(data, 1, n, f);
    uLong crc = crc32(0L, reinterpret_cast<const Bytef *>(type), 4);
    if (n) crc = crc32(crc, data, static_cast<uInt>(n));
    unsigned char c[4] = { (unsigned char)(crc >> 24), (unsigned char)(crc >> 16), (unsigned char)(crc >> 8), (unsigned char)crc };
    std::fwrite(c, 1, 4, f);
}
}  // namespace

extern "C" int bh_save_png(const char *path, const uint8_t *rgba8, uint32_t w, uint32_t h, int force_opaque)
{
    if (!path || !rgba8 || w == 0 || h == 0 || w > 65535 || h > 65535) { bh::set_error("bh_save_png: bad argument"); return BH_ERR_INVALID; }
    std::vector<unsigned char> raw((size_t)h * ((size_t)w * 4 + 1));
    for (uint32_t y = 0; y < h; ++y) {
        unsigned char *row = raw.data() + (size_t)y * ((size_t)w * 4 + 1);
        row[0] = 0;                                               // filter type None
        std::memcpy(row + 1, rgba8 + (size_t)y * w * 4, (size_t)w * 4);
        if (force_opaque) for (uint32_t x = 0; x < w; ++x) row[1 + 4 * (size_t)x + 3] = 255;
    }
    uLongf zn = compressBound(static_cast<uLong>(raw.size()));
    std::vector<unsigned char> z(zn);
    if (compress2(z.data(), &zn, raw.data(), static_cast<uLong>(raw.size()), 6) != Z_OK) { bh::set_error("bh_save_png: zlib failed"); return BH_ERR_NOMEM; }
    FILE *f = std::fopen(path, "wb");
    if (!f) { bh::set_error("bh_save_png: cannot open %s for writing", path); return BH_ERR_NOENT; }
    static const unsigned char sig[8] = { 0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a };
    std::fwrite(sig, 1, 8, f);
    unsigned char ihdr[13] = { (unsigned char)(w >> 24), (unsigned char)(w >> 16), (unsigned char)(w >> 8), (unsigned char)w,
                               (unsigned char)(h >> 24), (unsigned char)(h >> 16), (unsigned char)(h >> 8), (unsigned char)h,
                               8, 6, 0, 0, 0 };                  // 8-bit, colour type 6 (RGBA)
    png_chunk(f, "IHDR", ihdr, 13);
    png_chunk(f, "IDAT", z.data(), zn);
    png_chunk(f, "IEND", nullptr, 0);
    const bool ok = std::ferror(f) == 0;
    std::fclose(f);
    if (!ok) { bh::set_error("bh_save_png: write error on %s", path); return BH_ERR_CUDA; }
    return BH_OK;
}
