// hlm_netcdf4_writer.hpp — NetCDF-4 (HDF5) files without libnetcdf / libhdf5.
//
// The reference writes its outputs through netcdf-c as NC_NETCDF4 with nc_def_var_deflate(shuffle = 1,
// deflate = 1, level) on `outputs` (I_O/output_series.cpp:31,56,88,109).  Neither library exists in this image, so
// this header assembles the same container by hand, object for object as netcdf-c 4.9.2 / HDF5 1.14 laid out the
// reference's own committed files (src/final_example.nc, src/dense_example.nc — the structures below were read off
// their bytes):
//   * superblock version 2 (48 bytes, lookup3 checksum), root group object header at 48;
//   * version-2 object headers ("OHDR", creation order of attributes tracked and indexed, one chunk, checksum);
//   * root group: link info + group info + one compact hard-link message per variable, creation order = definition
//     order (netcdf-c orders variables by it), attribute _NCProperties;
//   * every dimension is a dimension scale: its coordinate variable carries CLASS = "DIMENSION_SCALE", NAME,
//     _Netcdf4Dimid and REFERENCE_LIST {dataset, dimension}; every variable carries _Netcdf4Coordinates (its dimids)
//     and, unless it is a coordinate variable, DIMENSION_LIST (variable-length lists of object references, stored in
//     a global heap collection);
//   * coordinate variables contiguous; `outputs` chunked (layout version 3, version-1 B-tree chunk index, 64-way
//     nodes) with the filter pipeline shuffle(element size) + deflate(level), or contiguous when level == 0 —
//     a variable that fits one 4 MiB chunk is ONE chunk, as in the reference's files;
//   * little-endian IEEE / two's-complement datatypes, netCDF default fill values.
// zlib's compress2() is the only dependency.  The reader of hlm_netcdf.hpp (Hdf5Reader) decodes these files with
// the code path that decodes the reference's own.
#pragma once

#include <zlib.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace hlmnc {

/// Bob Jenkins' lookup3 hashlittle() (public domain), initval 0: HDF5's metadata checksum (H5_checksum_lookup3).
inline uint32_t lookup3(const uint8_t* k, size_t length, uint32_t initval = 0) {
    auto rot = [](uint32_t x, int r) { return (x << r) | (x >> (32 - r)); };
    uint32_t a, b, c;
    a = b = c = 0xdeadbeefu + (uint32_t)length + initval;
    auto rd = [&](size_t i) { return (uint32_t)k[i] | ((uint32_t)k[i + 1] << 8) | ((uint32_t)k[i + 2] << 16) | ((uint32_t)k[i + 3] << 24); };
    while (length > 12) {
        a += rd(0); b += rd(4); c += rd(8);
        a -= c; a ^= rot(c, 4); c += b;
        b -= a; b ^= rot(a, 6); a += c;
        c -= b; c ^= rot(b, 8); b += a;
        a -= c; a ^= rot(c, 16); c += b;
        b -= a; b ^= rot(a, 19); a += c;
        c -= b; c ^= rot(b, 4); b += a;
        length -= 12;
        k += 12;
    }
    switch (length) {  // all the case statements fall through
    case 12: c += (uint32_t)k[11] << 24; /* fallthrough */
    case 11: c += (uint32_t)k[10] << 16; /* fallthrough */
    case 10: c += (uint32_t)k[9] << 8; /* fallthrough */
    case 9: c += k[8]; /* fallthrough */
    case 8: b += (uint32_t)k[7] << 24; /* fallthrough */
    case 7: b += (uint32_t)k[6] << 16; /* fallthrough */
    case 6: b += (uint32_t)k[5] << 8; /* fallthrough */
    case 5: b += k[4]; /* fallthrough */
    case 4: a += (uint32_t)k[3] << 24; /* fallthrough */
    case 3: a += (uint32_t)k[2] << 16; /* fallthrough */
    case 2: a += (uint32_t)k[1] << 8; /* fallthrough */
    case 1: a += k[0]; break;
    case 0: return c;
    }
    c ^= b; c -= rot(b, 14);
    a ^= c; a -= rot(c, 11);
    b ^= a; b -= rot(a, 25);
    c ^= b; c -= rot(b, 16);
    a ^= c; a -= rot(c, 4);
    b ^= a; b -= rot(a, 14);
    c ^= b; c -= rot(b, 24);
    return c;
}

class Nc4Writer {
  public:
    enum Type { kInt = 4, kFloat = 5, kDouble = 6 };  // the NC_INT / NC_FLOAT / NC_DOUBLE codes of hlm_netcdf.hpp
    static constexpr uint64_t kUndef = ~0ULL;
    static constexpr uint64_t kDefaultChunkBytes = 4ULL << 20;  // netcdf-c's DEFAULT_CHUNK_SIZE

    explicit Nc4Writer(const std::string& path) : path_(path) {
        f_ = std::fopen(path.c_str(), "wb");
        if (!f_) throw std::runtime_error("NetCDF-4: cannot create " + path);
        const uint8_t zero[48] = {0};
        write_at_end(zero, 48);  // the superblock is written last, when the root header's address is known
    }
    ~Nc4Writer() {
        if (f_) std::fclose(f_);
    }
    Nc4Writer(const Nc4Writer&) = delete;
    Nc4Writer& operator=(const Nc4Writer&) = delete;

    int def_dim(const std::string& name, uint64_t len) {
        dims_.push_back({name, len, -1});
        return (int)dims_.size() - 1;
    }
    int def_var(const std::string& name, int type, const std::vector<int>& dimids) {
        if (type != kInt && type != kFloat && type != kDouble) throw std::runtime_error("NetCDF-4 writer: int, float and double variables only");
        Var v;
        v.name = name;
        v.type = type;
        v.dimids = dimids;
        for (int d : dimids)
            if (d < 0 || d >= (int)dims_.size()) throw std::runtime_error("NetCDF-4 writer: bad dimension id");
        if (dimids.size() == 1 && dims_[(size_t)dimids[0]].name == name) dims_[(size_t)dimids[0]].scale_var = (int)vars_.size();
        vars_.push_back(std::move(v));
        return (int)vars_.size() - 1;
    }
    /// nc_def_var_deflate(ncid, varid, shuffle, deflate = level > 0, level): the variable becomes chunked.  chunk = {} picks
    /// the default: the whole variable if it fits 4 MiB, else slabs of the first dimension of at most 4 MiB.
    void def_var_deflate(int varid, bool shuffle, int level, std::vector<uint64_t> chunk = {}) {
        Var& v = vars_.at((size_t)varid);
        if (level <= 0 && chunk.empty()) return;  // output_series.cpp:56: no filter when the level is 0
        v.chunked = true;
        v.shuffle = shuffle && level > 0;
        v.deflate = level > 0 ? std::min(level, 9) : 0;
        if (chunk.empty()) {
            for (int d : v.dimids) chunk.push_back(std::max<uint64_t>(dims_[(size_t)d].len, 1));
            uint64_t row = elem_size(v.type);
            for (size_t k = 1; k < chunk.size(); ++k) row *= chunk[k];
            if (!chunk.empty() && row * chunk[0] > kDefaultChunkBytes) chunk[0] = std::max<uint64_t>(1, kDefaultChunkBytes / std::max<uint64_t>(row, 1));
        }
        if (chunk.size() != v.dimids.size()) throw std::runtime_error("NetCDF-4 writer: chunk rank differs from the variable's");
        v.chunk = chunk;
    }
    /// varid < 0: a global attribute
    void put_att_text(int varid, const std::string& name, const std::string& text) {
        (varid < 0 ? gatts_ : vars_.at((size_t)varid).atts).push_back({name, text});
    }

    /// the whole variable, row-major, in the machine's (little-endian) representation
    void put_var(int varid, const void* data) {
        Var& v = vars_.at((size_t)varid);
        const uint64_t es = elem_size(v.type);
        if (!v.chunked) {
            v.data_addr = align_end(8);
            v.data_size = count(v) * es;
            write_at_end(data, v.data_size);
            return;
        }
        // cut into chunks (edge chunks padded with the fill value), row-major over the chunk grid
        const size_t rank = v.dimids.size();
        std::vector<uint64_t> shape(rank), grid(rank), idx(rank, 0);
        uint64_t nchunks = 1, chunk_elems = 1;
        for (size_t k = 0; k < rank; ++k) {
            shape[k] = dims_[(size_t)v.dimids[k]].len;
            grid[k] = (shape[k] + v.chunk[k] - 1) / v.chunk[k];
            nchunks *= grid[k];
            chunk_elems *= v.chunk[k];
        }
        std::vector<uint8_t> buf(chunk_elems * es);
        for (uint64_t c = 0; c < nchunks; ++c) {
            fill_with_default(v.type, buf.data(), chunk_elems);
            copy_block(static_cast<const uint8_t*>(data), shape, v.chunk, idx, es, buf.data());
            put_chunk(varid, idx, buf.data());
            for (size_t k = rank; k-- > 0;) {
                if (++idx[k] < grid[k]) break;
                idx[k] = 0;
            }
        }
    }
    /// One full chunk (chunk-shaped, padded where it overhangs the variable) at chunk-grid position `chunk_index`.
    /// Chunks may arrive in any order, each once.  Thread-compatible with compress_chunk() done by the caller:
    /// see put_compressed_chunk.
    void put_chunk(int varid, const std::vector<uint64_t>& chunk_index, const void* chunk_data) {
        Var& v = vars_.at((size_t)varid);
        std::vector<uint8_t> out;
        encode_chunk(v, chunk_data, out);
        put_encoded_chunk(varid, chunk_index, out);
    }
    /// shuffle + deflate of one chunk as the variable's filter pipeline prescribes (callable from worker threads)
    void encode_chunk(int varid, const void* chunk_data, std::vector<uint8_t>& out) const { encode_chunk(vars_.at((size_t)varid), chunk_data, out); }
    /// append an encoded chunk and enter it into the variable's index (one thread at a time)
    void put_encoded_chunk(int varid, const std::vector<uint64_t>& chunk_index, const std::vector<uint8_t>& bytes) {
        Var& v = vars_.at((size_t)varid);
        if (!v.chunked) throw std::runtime_error("NetCDF-4 writer: put_chunk on a contiguous variable");
        ChunkRec r;
        r.offset.resize(v.chunk.size());
        for (size_t k = 0; k < v.chunk.size(); ++k) r.offset[k] = chunk_index[k] * v.chunk[k];
        r.addr = align_end(8);
        r.size = (uint32_t)bytes.size();
        write_at_end(bytes.data(), bytes.size());
        v.chunks.push_back(std::move(r));
    }
    uint64_t chunk_elems(int varid) const {
        uint64_t n = 1;
        for (uint64_t c : vars_.at((size_t)varid).chunk) n *= c;
        return n;
    }
    const std::vector<uint64_t>& chunk_shape(int varid) const { return vars_.at((size_t)varid).chunk; }
    static void fill_with_default(int type, void* dst, uint64_t n) {  // netcdf.h NC_FILL_INT / _FLOAT / _DOUBLE
        if (type == kInt) {
            const int32_t f = -2147483647;
            for (uint64_t i = 0; i < n; ++i) static_cast<int32_t*>(dst)[i] = f;
        } else if (type == kFloat) {
            const float f = 9.9692099683868690e+36f;
            for (uint64_t i = 0; i < n; ++i) static_cast<float*>(dst)[i] = f;
        } else {
            const double f = 9.9692099683868690e+36;
            for (uint64_t i = 0; i < n; ++i) static_cast<double*>(dst)[i] = f;
        }
    }
    static uint64_t elem_size(int type) { return type == kDouble ? 8 : 4; }

    void close() {
        if (!f_) return;
        for (const Dim& d : dims_)
            if (d.scale_var < 0) throw std::runtime_error("NetCDF-4 writer: dimension '" + d.name + "' needs a coordinate variable of the same name");
        for (Var& v : vars_) {
            if (!v.chunked && v.data_addr == kUndef) {  // never written: all fill values
                std::vector<uint8_t> buf(count(v) * elem_size(v.type));
                fill_with_default(v.type, buf.data(), count(v));
                v.data_addr = align_end(8);
                v.data_size = buf.size();
                write_at_end(buf.data(), buf.size());
            }
        }
        // ---- addresses of the metadata block: global heap, chunk B-trees, object headers ----
        uint64_t at = align_end(8);
        bool need_heap = false;
        for (size_t i = 0; i < vars_.size(); ++i) need_heap = need_heap || (!is_scale(i) && !vars_[i].dimids.empty());
        const uint64_t heap_addr = need_heap ? at : kUndef;
        std::vector<uint8_t> heap;
        if (need_heap) {
            heap = build_global_heap();  // sized now, filled with header addresses below
            at += heap.size();
        }
        std::vector<std::vector<uint8_t>> trees(vars_.size());
        for (size_t i = 0; i < vars_.size(); ++i)
            if (vars_[i].chunked) {
                vars_[i].btree_addr = at;
                trees[i] = build_chunk_btree(vars_[i], at);
                at += trees[i].size();
            }
        // object header sizes do not depend on the addresses inside them: size first, place, then serialise
        for (size_t i = 0; i < vars_.size(); ++i) {
            vars_[i].ohdr_addr = at;
            at += build_var_header(i, heap_addr).size();
        }
        const uint64_t root_addr = at;
        const std::vector<uint8_t> root = build_root_header();
        at += root.size();
        // ---- write ----
        if (need_heap) {
            heap = build_global_heap();
            write_at_end(heap.data(), heap.size());
        }
        for (size_t i = 0; i < vars_.size(); ++i)
            if (vars_[i].chunked) write_at_end(trees[i].data(), trees[i].size());
        for (size_t i = 0; i < vars_.size(); ++i) {
            const std::vector<uint8_t> h = build_var_header(i, heap_addr);
            if (end_ != vars_[i].ohdr_addr) throw std::logic_error("NetCDF-4 writer: header placement");
            write_at_end(h.data(), h.size());
        }
        write_at_end(root.data(), root.size());
        // superblock version 2
        std::vector<uint8_t> sb;
        const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
        sb.insert(sb.end(), sig, sig + 8);
        sb.push_back(2);  // version
        sb.push_back(8);  // size of offsets
        sb.push_back(8);  // size of lengths
        sb.push_back(0);  // file consistency flags
        p64(sb, 0);       // base address
        p64(sb, kUndef);  // superblock extension
        p64(sb, end_);    // end of file
        p64(sb, root_addr);
        p32(sb, lookup3(sb.data(), sb.size()));
        if (std::fseek(f_, 0, SEEK_SET) != 0 || std::fwrite(sb.data(), 1, sb.size(), f_) != sb.size())
            throw std::runtime_error("NetCDF-4: cannot write the superblock of " + path_);
        if (std::fclose(f_) != 0) {
            f_ = nullptr;
            throw std::runtime_error("NetCDF-4: closing " + path_ + " failed");
        }
        f_ = nullptr;
    }

  private:
    struct Att { std::string name, text; };
    struct Dim { std::string name; uint64_t len; int scale_var; };
    struct ChunkRec { std::vector<uint64_t> offset; uint64_t addr; uint32_t size; };
    struct Var {
        std::string name;
        int type = kDouble;
        std::vector<int> dimids;
        std::vector<Att> atts;
        bool chunked = false, shuffle = false;
        int deflate = 0;
        std::vector<uint64_t> chunk;
        uint64_t data_addr = kUndef, data_size = 0, btree_addr = kUndef, ohdr_addr = kUndef;
        std::vector<ChunkRec> chunks;
    };

    // ---- byte helpers -------------------------------------------------------------------------
    static void p16(std::vector<uint8_t>& b, uint32_t v) { b.push_back(v & 0xff); b.push_back((v >> 8) & 0xff); }
    static void p32(std::vector<uint8_t>& b, uint32_t v) { for (int i = 0; i < 4; ++i) b.push_back((v >> (8 * i)) & 0xff); }
    static void p64(std::vector<uint8_t>& b, uint64_t v) { for (int i = 0; i < 8; ++i) b.push_back((uint8_t)((v >> (8 * i)) & 0xff)); }
    static void pbytes(std::vector<uint8_t>& b, const void* p, size_t n) { b.insert(b.end(), static_cast<const uint8_t*>(p), static_cast<const uint8_t*>(p) + n); }
    static void pstr0(std::vector<uint8_t>& b, const std::string& s) { pbytes(b, s.data(), s.size()); b.push_back(0); }

    void write_at_end(const void* p, size_t n) {
        if (n && std::fwrite(p, 1, n, f_) != n) throw std::runtime_error("NetCDF-4: write to " + path_ + " failed (disk full?)");
        end_ += n;
    }
    uint64_t align_end(uint64_t a) {
        const uint8_t zero[8] = {0};
        const uint64_t pad = (a - end_ % a) % a;
        write_at_end(zero, pad);
        return end_;
    }
    uint64_t count(const Var& v) const {
        uint64_t n = 1;
        for (int d : v.dimids) n *= dims_[(size_t)d].len;
        return n;
    }
    bool is_scale(size_t var) const {
        const Var& v = vars_[var];
        return v.dimids.size() == 1 && dims_[(size_t)v.dimids[0]].scale_var == (int)var;
    }

    // one chunk-shaped block out of a row-major array, clipped at the array's edge
    static void copy_block(const uint8_t* src, const std::vector<uint64_t>& shape, const std::vector<uint64_t>& chunk,
                           const std::vector<uint64_t>& cidx, uint64_t es, uint8_t* dst) {
        const size_t rank = shape.size();
        if (rank == 0) { std::memcpy(dst, src, es); return; }
        std::vector<uint64_t> lo(rank), n(rank), pos(rank, 0);
        for (size_t k = 0; k < rank; ++k) {
            lo[k] = cidx[k] * chunk[k];
            n[k] = lo[k] >= shape[k] ? 0 : std::min(chunk[k], shape[k] - lo[k]);
            if (n[k] == 0) return;
        }
        const uint64_t run = n[rank - 1] * es;  // contiguous along the last dimension
        for (;;) {
            uint64_t s_off = 0, d_off = 0;
            for (size_t k = 0; k < rank; ++k) {
                const uint64_t p = k + 1 < rank ? pos[k] : 0;
                s_off = s_off * shape[k] + lo[k] + p;
                d_off = d_off * chunk[k] + p;
            }
            std::memcpy(dst + d_off * es, src + s_off * es, run);
            size_t k = rank - 1;  // next index tuple over dimensions 0 .. rank-2
            while (k > 0) {
                --k;
                if (++pos[k] < n[k]) break;
                pos[k] = 0;
                if (k == 0) return;
            }
            if (rank == 1) return;
        }
    }

    void encode_chunk(const Var& v, const void* chunk_data, std::vector<uint8_t>& out) const {
        uint64_t n = 1;
        for (uint64_t c : v.chunk) n *= c;
        const uint64_t es = elem_size(v.type), bytes = n * es;
        const uint8_t* src = static_cast<const uint8_t*>(chunk_data);
        std::vector<uint8_t> shuf;
        if (v.shuffle && es > 1) {  // HDF5 shuffle: byte j of every element together
            shuf.resize(bytes);
            for (uint64_t j = 0; j < es; ++j) {
                uint8_t* d = shuf.data() + j * n;
                for (uint64_t i = 0; i < n; ++i) d[i] = src[i * es + j];
            }
            src = shuf.data();
        }
        if (v.deflate > 0) {
            uLongf cap = compressBound((uLong)bytes);
            out.resize(cap);
            if (compress2(out.data(), &cap, src, (uLong)bytes, v.deflate) != Z_OK) throw std::runtime_error("NetCDF-4: deflate failed");
            out.resize(cap);
        } else {
            out.assign(src, src + bytes);
        }
    }

    // ---- datatype / dataspace encodings (as in the reference's files) ---------------------------------
    static std::vector<uint8_t> dt_of(int type) {
        std::vector<uint8_t> b;
        if (type == kInt) b = {0x10, 0x08, 0x00, 0x00, 4, 0, 0, 0, 0x00, 0x00, 0x20, 0x00};  // signed 32-bit LE
        else if (type == kFloat) b = {0x11, 0x20, 0x1f, 0x00, 4, 0, 0, 0, 0x00, 0x00, 0x20, 0x00, 0x17, 0x08, 0x00, 0x17, 0x7f, 0, 0, 0};
        else b = {0x11, 0x20, 0x3f, 0x00, 8, 0, 0, 0, 0x00, 0x00, 0x40, 0x00, 0x34, 0x0b, 0x00, 0x34, 0xff, 0x03, 0, 0};
        return b;
    }
    static std::vector<uint8_t> dt_uint32() { return {0x10, 0x00, 0x00, 0x00, 4, 0, 0, 0, 0x00, 0x00, 0x20, 0x00}; }
    static std::vector<uint8_t> dt_objref() { return {0x17, 0x00, 0x00, 0x00, 8, 0, 0, 0}; }
    static std::vector<uint8_t> dt_string(uint32_t n) {
        std::vector<uint8_t> b = {0x13, 0x00, 0x00, 0x00};
        p32(b, n);
        return b;
    }
    static std::vector<uint8_t> ds_scalar() { return {0x02, 0x00, 0x00, 0x00}; }
    static std::vector<uint8_t> ds_simple(const std::vector<uint64_t>& dims) {
        std::vector<uint8_t> b = {0x02, (uint8_t)dims.size(), 0x01, 0x01};  // version 2, rank, max dims present, simple
        for (uint64_t d : dims) p64(b, d);
        for (uint64_t d : dims) p64(b, d);
        return b;
    }
    // attribute message, version 3
    static std::vector<uint8_t> attr_msg(const std::string& name, const std::vector<uint8_t>& dt, const std::vector<uint8_t>& ds,
                                         const std::vector<uint8_t>& data) {
        std::vector<uint8_t> b = {0x03, 0x00};
        p16(b, (uint32_t)name.size() + 1);
        p16(b, (uint32_t)dt.size());
        p16(b, (uint32_t)ds.size());
        b.push_back(0);  // ASCII
        pstr0(b, name);
        pbytes(b, dt.data(), dt.size());
        pbytes(b, ds.data(), ds.size());
        pbytes(b, data.data(), data.size());
        return b;
    }
    static std::vector<uint8_t> attr_text(const std::string& name, const std::string& text, bool with_nul) {
        std::vector<uint8_t> data(text.begin(), text.end());
        if (with_nul || data.empty()) data.push_back(0);
        return attr_msg(name, dt_string((uint32_t)data.size()), ds_scalar(), data);
    }
    static std::vector<uint8_t> attr_ints(const std::string& name, const std::vector<int32_t>& v, bool scalar) {
        std::vector<uint8_t> data;
        for (int32_t x : v) p32(data, (uint32_t)x);
        return attr_msg(name, dt_of(kInt), scalar ? ds_scalar() : ds_simple({v.size()}), data);
    }

    // ---- version-2 object header: one chunk, creation order of attributes tracked + indexed ----------
    struct Msg { uint8_t type, flags; uint16_t corder; std::vector<uint8_t> body; };
    static std::vector<uint8_t> ohdr(const std::vector<Msg>& msgs) {
        uint64_t body = 0;
        for (const Msg& m : msgs) body += 6 + m.body.size();
        std::vector<uint8_t> b = {'O', 'H', 'D', 'R', 2, 0x0e};  // flags: 4-byte chunk size | attr creation order tracked | indexed
        p32(b, (uint32_t)body);
        for (const Msg& m : msgs) {
            b.push_back(m.type);
            p16(b, (uint32_t)m.body.size());
            b.push_back(m.flags);
            p16(b, m.corder);
            pbytes(b, m.body.data(), m.body.size());
        }
        p32(b, lookup3(b.data(), b.size()));
        return b;
    }
    static std::vector<uint8_t> attr_info(uint16_t max_corder) {
        std::vector<uint8_t> b = {0x00, 0x03};  // version 0; creation order tracked + indexed
        p16(b, max_corder);
        p64(b, kUndef);  // fractal heap (dense storage): none, attributes are compact
        p64(b, kUndef);  // name index
        p64(b, kUndef);  // creation order index
        return b;
    }

    // the (variable, dimension index) pairs that use dimension d, in variable order: REFERENCE_LIST of its scale
    std::vector<std::pair<size_t, uint32_t>> users_of(int d) const {
        std::vector<std::pair<size_t, uint32_t>> u;
        for (size_t i = 0; i < vars_.size(); ++i) {
            if (is_scale(i)) continue;
            for (size_t k = 0; k < vars_[i].dimids.size(); ++k)
                if (vars_[i].dimids[k] == d) u.push_back({i, (uint32_t)k});
        }
        return u;
    }
    // global heap object ids: object (1 + position) per (non-scale variable, dimension), in variable order
    uint32_t heap_index(size_t var, size_t k) const {
        uint32_t id = 1;
        for (size_t i = 0; i < vars_.size(); ++i) {
            if (is_scale(i)) continue;
            for (size_t j = 0; j < vars_[i].dimids.size(); ++j) {
                if (i == var && j == k) return id;
                ++id;
            }
        }
        return 0;
    }
    std::vector<uint8_t> build_global_heap() const {
        std::vector<uint8_t> b = {'G', 'C', 'O', 'L', 1, 0, 0, 0};
        const size_t size_at = b.size();
        p64(b, 0);  // collection size, patched
        for (size_t i = 0; i < vars_.size(); ++i) {
            if (is_scale(i)) continue;
            for (size_t k = 0; k < vars_[i].dimids.size(); ++k) {
                p16(b, heap_index(i, k));
                p16(b, 1);  // reference count
                p32(b, 0);
                p64(b, 8);  // object size: one object reference
                const int sv = dims_[(size_t)vars_[i].dimids[k]].scale_var;
                p64(b, vars_[(size_t)sv].ohdr_addr);
            }
        }
        const uint64_t total = std::max<uint64_t>(4096, (b.size() + 16 + 7) / 8 * 8);  // minimum collection size 4096
        p16(b, 0);  // object 0: the free space
        p16(b, 0);
        p32(b, 0);
        p64(b, total - b.size() - 8);
        b.resize(total, 0);
        for (int i = 0; i < 8; ++i) b[size_at + i] = (uint8_t)((total >> (8 * i)) & 0xff);
        return b;
    }

    // version-1 B-tree of raw-data chunks (node type 1), 64-way nodes (istore_k = 32, HDF5's default)
    static constexpr int kNodeEntries = 64;
    std::vector<uint8_t> build_chunk_btree(Var& v, uint64_t base) const {
        const size_t rank = v.chunk.size();
        std::sort(v.chunks.begin(), v.chunks.end(), [](const ChunkRec& a, const ChunkRec& b) { return a.offset < b.offset; });
        const uint64_t key_size = 8 + 8 * (rank + 1);
        const uint64_t node_size = 24 + (kNodeEntries + 1) * key_size + kNodeEntries * 8;
        struct Ent { std::vector<uint64_t> offset; uint32_t size; uint64_t child; };
        std::vector<Ent> level;
        for (const ChunkRec& c : v.chunks) level.push_back({c.offset, c.size, c.addr});
        std::vector<uint64_t> past(rank);  // the key after the last chunk: one chunk beyond the variable's extent
        for (size_t k = 0; k < rank; ++k) past[k] = dims_[(size_t)v.dimids[k]].len;
        std::vector<uint8_t> out;
        auto put_key = [&](std::vector<uint8_t>& b, uint32_t size, const std::vector<uint64_t>& off, uint64_t last) {
            p32(b, size);
            p32(b, 0);  // filter mask: every filter applied
            for (uint64_t o : off) p64(b, o);
            p64(b, last);  // the extra dimension of the element size
        };
        if (level.empty()) level.push_back({std::vector<uint64_t>(rank, 0), 0, kUndef});  // no chunk written: an empty leaf
        int depth = 0;
        for (;;) {
            std::vector<Ent> parents;
            const size_t n_nodes = (level.size() + kNodeEntries - 1) / kNodeEntries;
            const uint64_t level_base = base + out.size();
            for (size_t n = 0; n < n_nodes; ++n) {
                const size_t lo = n * kNodeEntries, hi = std::min(level.size(), lo + kNodeEntries);
                std::vector<uint8_t> b = {'T', 'R', 'E', 'E', 1, (uint8_t)depth};
                const bool empty = level[lo].child == kUndef;
                p16(b, empty ? 0 : (uint32_t)(hi - lo));
                p64(b, n > 0 ? level_base + (n - 1) * node_size : kUndef);          // left sibling
                p64(b, n + 1 < n_nodes ? level_base + (n + 1) * node_size : kUndef);  // right sibling
                for (size_t e = lo; e < hi && !empty; ++e) {
                    put_key(b, level[e].size, level[e].offset, 0);
                    p64(b, level[e].child);
                }
                if (hi < level.size()) put_key(b, level[hi].size, level[hi].offset, 0);  // the next node's first key
                else put_key(b, 0, past, elem_size(v.type));
                b.resize(node_size, 0);
                parents.push_back({level[lo].offset, level[lo].size, level_base + n * node_size});
                out.insert(out.end(), b.begin(), b.end());
            }
            if (n_nodes == 1) {
                v.btree_addr = level_base;  // the root is the last node written
                break;
            }
            level = std::move(parents);
            ++depth;
        }
        return out;
    }

    std::vector<uint8_t> build_var_header(size_t i, uint64_t heap_addr) const {
        const Var& v = vars_[i];
        std::vector<Msg> m;
        std::vector<uint64_t> shape;
        for (int d : v.dimids) shape.push_back(dims_[(size_t)d].len);
        m.push_back({0x01, 0, 0, shape.empty() ? ds_scalar() : ds_simple(shape)});
        m.push_back({0x03, 1, 0, dt_of(v.type)});
        {   // fill value, version 3: allocation late (contiguous) / incremental (chunked), written if set, defined
            std::vector<uint8_t> b = {0x03, (uint8_t)(v.chunked ? 0x2b : 0x2a)};
            p32(b, (uint32_t)elem_size(v.type));
            uint8_t fv[8];
            fill_with_default(v.type, fv, 1);
            pbytes(b, fv, elem_size(v.type));
            m.push_back({0x05, 1, 0, b});
        }
        if (v.chunked && v.deflate > 0) {  // filter pipeline, version 2: shuffle(element size), deflate(level)
            std::vector<uint8_t> b = {0x02, (uint8_t)(v.shuffle ? 2 : 1)};
            if (v.shuffle) { p16(b, 2); p16(b, 1); p16(b, 1); p32(b, (uint32_t)elem_size(v.type)); }
            p16(b, 1); p16(b, 1); p16(b, 1); p32(b, (uint32_t)v.deflate);
            m.push_back({0x0b, 1, 0, b});
        }
        {   // data layout, version 3
            std::vector<uint8_t> b = {0x03, (uint8_t)(v.chunked ? 2 : 1)};
            if (v.chunked) {
                b.push_back((uint8_t)(v.chunk.size() + 1));
                p64(b, v.btree_addr);
                for (uint64_t c : v.chunk) p32(b, (uint32_t)c);
                p32(b, (uint32_t)elem_size(v.type));
            } else {
                p64(b, v.data_addr);
                p64(b, v.data_size);
            }
            m.push_back({0x08, 0, 0, b});
        }
        // attributes, creation order as netcdf-c writes them
        std::vector<std::vector<uint8_t>> atts;
        {
            std::vector<int32_t> ids(v.dimids.begin(), v.dimids.end());
            atts.push_back(attr_ints("_Netcdf4Coordinates", ids, false));
        }
        if (is_scale(i)) {
            atts.push_back(attr_text("CLASS", "DIMENSION_SCALE", true));
            atts.push_back(attr_text("NAME", v.name, true));
            atts.push_back(attr_ints("_Netcdf4Dimid", {v.dimids[0]}, true));
        }
        for (const Att& a : v.atts) atts.push_back(attr_text(a.name, a.text, false));
        if (is_scale(i)) {
            const auto users = users_of(v.dimids[0]);
            if (!users.empty()) {  // REFERENCE_LIST: compound {dataset: object reference @0, dimension: uint32 @8}, 16 bytes
                std::vector<uint8_t> dt = {0x36, 0x02, 0x00, 0x00};
                p32(dt, 16);
                pstr0(dt, "dataset");
                dt.push_back(0);
                { const auto r = dt_objref(); pbytes(dt, r.data(), r.size()); }
                pstr0(dt, "dimension");
                dt.push_back(8);
                { const auto u = dt_uint32(); pbytes(dt, u.data(), u.size()); }
                std::vector<uint8_t> data;
                for (const auto& u : users) {
                    p64(data, vars_[u.first].ohdr_addr);
                    p32(data, u.second);
                    p32(data, 0);
                }
                atts.push_back(attr_msg("REFERENCE_LIST", dt, ds_simple({users.size()}), data));
            }
        } else if (!v.dimids.empty()) {  // DIMENSION_LIST: per dimension a variable-length list of one object reference
            std::vector<uint8_t> dt = {0x19, 0x00, 0x00, 0x00};
            p32(dt, 16);
            { const auto r = dt_objref(); pbytes(dt, r.data(), r.size()); }
            std::vector<uint8_t> data;
            for (size_t k = 0; k < v.dimids.size(); ++k) {
                p32(data, 1);
                p64(data, heap_addr);
                p32(data, heap_index(i, k));
            }
            atts.push_back(attr_msg("DIMENSION_LIST", dt, ds_simple({v.dimids.size()}), data));
        }
        m.push_back({0x15, 4, 0, attr_info((uint16_t)atts.size())});
        for (size_t a = 0; a < atts.size(); ++a) m.push_back({0x0c, 0, (uint16_t)a, atts[a]});
        return ohdr(m);
    }

    std::vector<uint8_t> build_root_header() const {
        std::vector<Msg> m;
        {   // link info, version 0: creation order tracked + indexed, compact storage
            std::vector<uint8_t> b = {0x00, 0x03};
            p64(b, vars_.size());
            p64(b, kUndef);
            p64(b, kUndef);
            p64(b, kUndef);
            m.push_back({0x02, 0, 0, b});
        }
        m.push_back({0x0a, 1, 0, {0x00, 0x00}});  // group info
        for (size_t i = 0; i < vars_.size(); ++i) {  // hard links, creation order = definition order
            std::vector<uint8_t> b = {0x01, 0x04};
            p64(b, i);
            if (vars_[i].name.size() > 255) throw std::runtime_error("NetCDF-4 writer: variable name too long");
            b.push_back((uint8_t)vars_[i].name.size());
            pbytes(b, vars_[i].name.data(), vars_[i].name.size());
            p64(b, vars_[i].ohdr_addr);
            m.push_back({0x06, 0, 0, b});
        }
        std::vector<std::vector<uint8_t>> atts;
        atts.push_back(attr_text("_NCProperties", "version=2,hlm_b200=2", false));
        for (const Att& a : gatts_) atts.push_back(attr_text(a.name, a.text, false));
        m.push_back({0x15, 4, 0, attr_info((uint16_t)atts.size())});
        for (size_t a = 0; a < atts.size(); ++a) m.push_back({0x0c, 0, (uint16_t)a, atts[a]});
        return ohdr(m);
    }

    std::string path_;
    std::FILE* f_ = nullptr;
    uint64_t end_ = 0;
    std::vector<Dim> dims_;
    std::vector<Var> vars_;
    std::vector<Att> gatts_;
};

}  // namespace hlmnc
