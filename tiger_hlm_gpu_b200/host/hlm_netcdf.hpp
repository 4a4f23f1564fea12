// hlm_netcdf.hpp — NetCDF input and output of the path, without libnetcdf.
//
// The reference reads gridded forcings with netcdf-c (I_O/forcing_loader.cpp:67-218, class
// NetCDFLoader) and writes its results with nc_create(NC_NETCDF4) (I_O/output_series.cpp:18-124).
// netcdf-c / HDF5 are not part of this image, and a solver library should not drag them in, so the
// two on-disk contracts are implemented here directly:
//
//   reading   * NetCDF classic (CDF-1, CDF-2 "64-bit offset", CDF-5), fixed and record variables;
//             * NetCDF-4 = HDF5, the subset netcdf-c itself writes for (time, lat, lon) grids and
//               that the reference's own outputs use (src/final_example.nc, src/dense_example.nc):
//               superblock v0-v3, object headers v1/v2, groups as symbol tables or compact link
//               messages, contiguous / compact / chunked (v1 B-tree) layouts, deflate + shuffle +
//               fletcher32 filters, fixed-point and IEEE datatypes of either byte order.
//             Values are converted to the requested type like nc_get_vara_float does (no
//             scale_factor / add_offset unpacking — the reference applies none either).
//   writing   NetCDF classic CDF-2 with the reference's dimension, variable and attribute names
//             (outputs(system,time,variable) / outputs(system,variable)).  Every netCDF tool reads
//             it; what is lost against NC_NETCDF4 is only the optional deflate filter.
//
// Header-only; needs zlib for deflate-compressed HDF5 chunks (-lz).
#pragma once

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <functional>
#include <cstdio>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

static_assert(__BYTE_ORDER__ == __ORDER_LITTLE_ENDIAN__, "host byte order assumed little-endian");

#include "hlm_netcdf4_writer.hpp"

namespace hlmnc {

enum NcType : int {
    NC_BYTE = 1, NC_CHAR = 2, NC_SHORT = 3, NC_INT = 4, NC_FLOAT = 5, NC_DOUBLE = 6,
    NC_UBYTE = 7, NC_USHORT = 8, NC_UINT = 9, NC_INT64 = 10, NC_UINT64 = 11
};

inline int nc_type_size(int t) {
    switch (t) {
        case NC_BYTE: case NC_CHAR: case NC_UBYTE: return 1;
        case NC_SHORT: case NC_USHORT: return 2;
        case NC_INT: case NC_UINT: case NC_FLOAT: return 4;
        case NC_DOUBLE: case NC_INT64: case NC_UINT64: return 8;
    }
    throw std::runtime_error("NetCDF: unknown nc_type " + std::to_string(t));
}

// One stored number format: how to turn `size` raw bytes into a value.
struct ElemType {
    enum Kind { kInt, kUInt, kFloat } kind = kFloat;
    int size = 4;
    bool big_endian = false;
};

inline ElemType elem_of_nc_type(int t) {
    ElemType e;
    e.big_endian = true;  // classic files are big-endian throughout
    e.size = nc_type_size(t);
    e.kind = (t == NC_FLOAT || t == NC_DOUBLE) ? ElemType::kFloat
             : (t == NC_UBYTE || t == NC_USHORT || t == NC_UINT || t == NC_UINT64 || t == NC_CHAR) ? ElemType::kUInt
                                                                                                  : ElemType::kInt;
    return e;
}

template <typename T> inline T load_elem(const uint8_t* p, const ElemType& e) {
    uint8_t b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (e.big_endian) for (int i = 0; i < e.size; ++i) b[i] = p[e.size - 1 - i];
    else std::memcpy(b, p, e.size);
    if (e.kind == ElemType::kFloat) {
        if (e.size == 4) { float v; std::memcpy(&v, b, 4); return (T)v; }
        double v; std::memcpy(&v, b, 8); return (T)v;
    }
    if (e.kind == ElemType::kUInt) {
        uint64_t v = 0; std::memcpy(&v, b, e.size); return (T)v;
    }
    switch (e.size) {
        case 1: { int8_t v; std::memcpy(&v, b, 1); return (T)v; }
        case 2: { int16_t v; std::memcpy(&v, b, 2); return (T)v; }
        case 4: { int32_t v; std::memcpy(&v, b, 4); return (T)v; }
        default: { int64_t v; std::memcpy(&v, b, 8); return (T)v; }
    }
}

struct Attribute {
    std::string name;
    bool is_text = false;
    std::string text;             // when is_text
    std::vector<double> numbers;  // otherwise
};

struct VarInfo {
    std::string name;
    std::vector<uint64_t> shape;
    std::vector<std::string> dim_names;  // empty strings when the container does not name them
    ElemType elem;
    std::vector<Attribute> atts;
    const Attribute* att(const std::string& n) const {
        for (auto& a : atts) if (a.name == n) return &a;
        return nullptr;
    }
};

// Read-only memory map of a whole file.
class MappedFile {
  public:
    explicit MappedFile(const std::string& path) : path_(path) {
        fd_ = ::open(path.c_str(), O_RDONLY);
        if (fd_ < 0) throw std::runtime_error("Opening file " + path + ": " + std::strerror(errno));
        struct stat st;
        if (fstat(fd_, &st) != 0) { ::close(fd_); throw std::runtime_error("stat " + path); }
        size_ = (uint64_t)st.st_size;
        if (size_ > 0) {
            void* p = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
            if (p == MAP_FAILED) { ::close(fd_); throw std::runtime_error("mmap " + path); }
            data_ = (const uint8_t*)p;
        }
    }
    ~MappedFile() {
        if (data_) munmap((void*)data_, size_);
        if (fd_ >= 0) ::close(fd_);
    }
    MappedFile(const MappedFile&) = delete;
    MappedFile& operator=(const MappedFile&) = delete;
    const uint8_t* at(uint64_t off, uint64_t n) const {
        if (off > size_ || n > size_ - off) throw std::runtime_error("NetCDF: read past end of " + path_);
        return data_ + off;
    }
    uint64_t size() const { return size_; }
    const std::string& path() const { return path_; }

  private:
    std::string path_;
    int fd_ = -1;
    const uint8_t* data_ = nullptr;
    uint64_t size_ = 0;
};

// Common reader interface.
class Reader {
  public:
    virtual ~Reader() {}
    virtual std::vector<std::string> variables() const = 0;
    virtual bool has_variable(const std::string& name) const = 0;
    virtual const VarInfo& inquire(const std::string& name) const = 0;
    /// Hyperslab start/count (one entry per dimension) converted to T, C order.
    virtual void read(const std::string& name, const std::vector<uint64_t>& start, const std::vector<uint64_t>& count,
                      const std::function<void(uint64_t dst_index, const uint8_t* src, uint64_t n_contig, const ElemType&)>& sink) const = 0;

    template <typename T>
    std::vector<T> read_as(const std::string& name, const std::vector<uint64_t>& start, const std::vector<uint64_t>& count) const {
        uint64_t n = 1;
        for (auto c : count) n *= c;
        std::vector<T> out(n);
        read_into<T>(name, start, count, out.data());
        return out;
    }
    template <typename T>
    void read_into(const std::string& name, const std::vector<uint64_t>& start, const std::vector<uint64_t>& count, T* out) const {
        read(name, start, count, [out](uint64_t dst, const uint8_t* src, uint64_t n, const ElemType& e) {
            for (uint64_t i = 0; i < n; ++i) out[dst + i] = load_elem<T>(src + i * e.size, e);
        });
    }
    template <typename T> std::vector<T> read_all(const std::string& name) const {
        const VarInfo& v = inquire(name);
        return read_as<T>(name, std::vector<uint64_t>(v.shape.size(), 0), v.shape);
    }
};

inline void check_slab(const VarInfo& v, const std::vector<uint64_t>& start, const std::vector<uint64_t>& count) {
    if (start.size() != v.shape.size() || count.size() != v.shape.size())
        throw std::runtime_error("NetCDF: hyperslab rank does not match variable " + v.name);
    for (size_t d = 0; d < v.shape.size(); ++d)
        if (start[d] > v.shape[d] || count[d] > v.shape[d] - start[d])
            throw std::out_of_range("NetCDF: hyperslab exceeds variable " + v.name);
}

// ---------------------------------------------------------------------------------------------
// NetCDF classic: CDF-1 / CDF-2 / CDF-5
// ---------------------------------------------------------------------------------------------
class ClassicReader : public Reader {
  public:
    explicit ClassicReader(std::shared_ptr<MappedFile> f) : f_(std::move(f)) { parse(); }

    std::vector<std::string> variables() const override {
        std::vector<std::string> n;
        for (auto& v : vars_) n.push_back(v.info.name);
        return n;
    }
    bool has_variable(const std::string& name) const override { return index_.count(name) != 0; }
    const VarInfo& inquire(const std::string& name) const override { return var(name).info; }

    void read(const std::string& name, const std::vector<uint64_t>& start, const std::vector<uint64_t>& count,
              const std::function<void(uint64_t, const uint8_t*, uint64_t, const ElemType&)>& sink) const override {
        const V& v = var(name);
        check_slab(v.info, start, count);
        const size_t nd = v.info.shape.size();
        const int es = v.info.elem.size;
        if (nd == 0) { sink(0, f_->at(v.begin, es), 1, v.info.elem); return; }
        uint64_t total = 1;
        for (auto c : count) total *= c;
        if (total == 0) return;
        if (v.record && nd == 1) {  // one element per record
            for (uint64_t i = 0; i < count[0]; ++i) sink(i, f_->at(v.begin + (start[0] + i) * recsize_, es), 1, v.info.elem);
            return;
        }
        // strides in elements inside one record (record variables: dimension 0 steps by recsize bytes)
        std::vector<uint64_t> stride(nd, 1);
        for (int d = (int)nd - 2; d >= 0; --d) stride[d] = stride[d + 1] * v.info.shape[d + 1];
        const uint64_t run = count[nd - 1];
        std::vector<uint64_t> idx(nd, 0);
        uint64_t dst = 0;
        for (;;) {
            uint64_t off = v.begin;
            for (size_t d = 0; d < nd; ++d) {
                const uint64_t i = start[d] + idx[d];
                if (d == 0 && v.record) off += i * recsize_;
                else off += i * stride[d] * es;
            }
            sink(dst, f_->at(off, run * es), run, v.info.elem);
            dst += run;
            int d = (int)nd - 2;
            for (; d >= 0; --d) {
                if (++idx[d] < count[d]) break;
                idx[d] = 0;
            }
            if (d < 0) break;
        }
    }
    int version() const { return version_; }
    uint64_t numrecs() const { return numrecs_; }

  private:
    struct V {
        VarInfo info;
        bool record = false;
        uint64_t vsize = 0, begin = 0;
    };
    const V& var(const std::string& name) const {
        auto it = index_.find(name);
        if (it == index_.end()) throw std::runtime_error("Variable " + name + " not found in file");
        return vars_[it->second];
    }
    uint64_t pos_ = 0;
    uint32_t u32() { const uint8_t* p = f_->at(pos_, 4); pos_ += 4; return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }
    uint64_t u64() { const uint64_t hi = u32(); return hi << 32 | u32(); }
    uint64_t count_field() { return version_ == 5 ? u64() : u32(); }
    std::string name() {
        const uint64_t n = count_field();
        const uint8_t* p = f_->at(pos_, n);
        pos_ += (n + 3) & ~uint64_t(3);
        return std::string((const char*)p, n);
    }
    std::vector<Attribute> att_list() {
        const uint32_t tag = u32();
        const uint64_t n = count_field();
        std::vector<Attribute> out;
        if (tag == 0 && n == 0) return out;
        if (tag != 0x0C) throw std::runtime_error("NetCDF classic: bad attribute list tag in " + f_->path());
        for (uint64_t i = 0; i < n; ++i) {
            Attribute a;
            a.name = name();
            const int t = (int)u32();
            const uint64_t ne = count_field();
            const int es = nc_type_size(t);
            const uint8_t* p = f_->at(pos_, ne * es);
            pos_ += (ne * es + 3) & ~uint64_t(3);
            if (t == NC_CHAR) {
                a.is_text = true;
                a.text.assign((const char*)p, ne);
            } else {
                const ElemType e = elem_of_nc_type(t);
                for (uint64_t k = 0; k < ne; ++k) a.numbers.push_back(load_elem<double>(p + k * es, e));
            }
            out.push_back(std::move(a));
        }
        return out;
    }
    void parse() {
        const uint8_t* m = f_->at(0, 4);
        if (m[0] != 'C' || m[1] != 'D' || m[2] != 'F' || (m[3] != 1 && m[3] != 2 && m[3] != 5))
            throw std::runtime_error("not a NetCDF classic file: " + f_->path());
        version_ = m[3];
        pos_ = 4;
        numrecs_ = count_field();
        const bool streaming = (version_ == 5) ? numrecs_ == ~uint64_t(0) : numrecs_ == 0xFFFFFFFFu;
        // dimensions
        std::vector<std::pair<std::string, uint64_t>> dims;
        {
            const uint32_t tag = u32();
            const uint64_t n = count_field();
            if (!(tag == 0 && n == 0)) {
                if (tag != 0x0A) throw std::runtime_error("NetCDF classic: bad dimension list tag in " + f_->path());
                for (uint64_t i = 0; i < n; ++i) {
                    std::string nm = name();
                    dims.emplace_back(nm, count_field());
                }
            }
        }
        att_list();  // global attributes: not needed
        {
            const uint32_t tag = u32();
            const uint64_t n = count_field();
            if (!(tag == 0 && n == 0)) {
                if (tag != 0x0B) throw std::runtime_error("NetCDF classic: bad variable list tag in " + f_->path());
                for (uint64_t i = 0; i < n; ++i) {
                    V v;
                    v.info.name = name();
                    const uint64_t nd = count_field();
                    std::vector<uint64_t> dimids(nd);
                    for (auto& d : dimids) d = count_field();
                    v.info.atts = att_list();
                    const int t = (int)u32();
                    v.info.elem = elem_of_nc_type(t);
                    v.vsize = count_field();
                    v.begin = (version_ == 1) ? u32() : u64();
                    for (size_t k = 0; k < nd; ++k) {
                        if (dimids[k] >= dims.size()) throw std::runtime_error("NetCDF classic: bad dimension id");
                        v.info.dim_names.push_back(dims[dimids[k]].first);
                        v.info.shape.push_back(dims[dimids[k]].second);
                        if (k == 0 && dims[dimids[k]].second == 0) v.record = true;  // the UNLIMITED dimension has length 0
                    }
                    vars_.push_back(std::move(v));
                }
            }
        }
        // record size: sum of the (padded) per-record sizes; a single record variable is not padded
        int n_rec = 0;
        uint64_t first_rec_begin = ~uint64_t(0);
        for (auto& v : vars_)
            if (v.record) {
                ++n_rec;
                uint64_t per = v.info.elem.size;
                for (size_t k = 1; k < v.info.shape.size(); ++k) per *= v.info.shape[k];
                v.vsize = per;
                first_rec_begin = std::min(first_rec_begin, v.begin);
            }
        recsize_ = 0;
        for (auto& v : vars_)
            if (v.record) recsize_ += (n_rec == 1) ? v.vsize : ((v.vsize + 3) & ~uint64_t(3));
        if (streaming && n_rec > 0 && recsize_ > 0) numrecs_ = (f_->size() - first_rec_begin) / recsize_;
        for (auto& v : vars_)
            if (v.record) v.info.shape[0] = numrecs_;
        for (size_t i = 0; i < vars_.size(); ++i) index_[vars_[i].info.name] = i;
    }

    std::shared_ptr<MappedFile> f_;
    int version_ = 1;
    uint64_t numrecs_ = 0, recsize_ = 0;
    std::vector<V> vars_;
    std::map<std::string, size_t> index_;
};

// ---------------------------------------------------------------------------------------------
// HDF5 (the NetCDF-4 container), read-only subset
// ---------------------------------------------------------------------------------------------
class Hdf5Reader : public Reader {
  public:
    explicit Hdf5Reader(std::shared_ptr<MappedFile> f) : f_(std::move(f)) { parse(); }

    std::vector<std::string> variables() const override {
        std::vector<std::string> n;
        for (auto& d : sets_) n.push_back(d.info.name);
        return n;
    }
    bool has_variable(const std::string& name) const override { return index_.count(name) != 0; }
    const VarInfo& inquire(const std::string& name) const override { return set(name).info; }

    void read(const std::string& name, const std::vector<uint64_t>& start, const std::vector<uint64_t>& count,
              const std::function<void(uint64_t, const uint8_t*, uint64_t, const ElemType&)>& sink) const override {
        const Dataset& d = set(name);
        check_slab(d.info, start, count);
        const size_t nd = d.info.shape.size();
        const int es = d.info.elem.size;
        uint64_t total = 1;
        for (auto c : count) total *= c;
        if (total == 0) return;
        if (nd == 0) {
            if (d.layout == 2) throw std::runtime_error("HDF5: chunked scalar dataset " + name);
            sink(0, d.layout == 0 ? d.compact.data() : f_->at(d.addr, es), 1, d.info.elem);
            return;
        }
        std::vector<uint64_t> dst_stride(nd, 1);
        for (int k = (int)nd - 2; k >= 0; --k) dst_stride[k] = dst_stride[k + 1] * count[k + 1];
        if (d.layout != 2) {
            // contiguous or compact: one "chunk" covering the whole dataset
            const uint8_t* base = nullptr;
            uint64_t nbytes = es;
            for (auto s : d.info.shape) nbytes *= s;
            std::vector<uint8_t> zeros;
            if (d.layout == 0) base = d.compact.data();
            else if (d.addr == kUndef) { zeros.assign(nbytes, 0); base = zeros.data(); }  // never written: fill value
            else base = f_->at(d.addr, nbytes);
            copy_block(base, std::vector<uint64_t>(nd, 0), d.info.shape, start, count, dst_stride, d.info.elem, sink);
            return;
        }
        // chunked: every element not covered by a stored chunk keeps the fill value 0
        {
            std::vector<uint8_t> z((size_t)count[nd - 1] * es, 0);
            std::vector<uint64_t> idx(nd, 0);
            uint64_t dst = 0;
            for (;;) {
                sink(dst, z.data(), count[nd - 1], d.info.elem);
                dst += count[nd - 1];
                int k = (int)nd - 2;
                for (; k >= 0; --k) {
                    if (++idx[k] < count[k]) break;
                    idx[k] = 0;
                }
                if (k < 0) break;
            }
        }
        std::vector<uint8_t> buf, tmp;
        uint64_t chunk_bytes = es;
        for (auto c : d.chunk) chunk_bytes *= c;
        walk_chunks(d, d.addr, [&](const std::vector<uint64_t>& off, uint64_t addr, uint32_t size, uint32_t mask) {
            for (size_t k = 0; k < nd; ++k)  // skip chunks outside the slab
                if (off[k] >= start[k] + count[k] || off[k] + d.chunk[k] <= start[k]) return;
            const uint8_t* raw = f_->at(addr, size);
            const uint8_t* data = raw;
            uint64_t len = size;
            // filters are undone in reverse order of the pipeline
            for (int fi = (int)d.filters.size() - 1; fi >= 0; --fi) {
                if (mask & (1u << fi)) continue;
                const int id = d.filters[fi].id;
                if (id == 3) {  // fletcher32: checksum trailer
                    if (len < 4) throw std::runtime_error("HDF5: short fletcher32 chunk");
                    len -= 4;
                } else if (id == 1) {  // deflate
                    tmp.resize(chunk_bytes + 16);
                    uLongf out_len = tmp.size();
                    int rc = uncompress(tmp.data(), &out_len, data, len);
                    if (rc == Z_BUF_ERROR) {
                        tmp.resize(chunk_bytes * 4 + 1024);
                        out_len = tmp.size();
                        rc = uncompress(tmp.data(), &out_len, data, len);
                    }
                    if (rc != Z_OK) throw std::runtime_error("HDF5: inflate failed for a chunk of " + name);
                    buf.swap(tmp);
                    data = buf.data();
                    len = out_len;
                } else if (id == 2) {  // shuffle: byte planes -> elements
                    const uint64_t ne = len / es;
                    tmp.resize(len);
                    for (int b = 0; b < es; ++b)
                        for (uint64_t e = 0; e < ne; ++e) tmp[e * es + b] = data[(uint64_t)b * ne + e];
                    std::memcpy(tmp.data() + ne * es, data + ne * es, len - ne * es);
                    buf.swap(tmp);
                    data = buf.data();
                } else {
                    throw std::runtime_error("HDF5: unsupported filter id " + std::to_string(id) + " on " + name);
                }
            }
            if (len < chunk_bytes) throw std::runtime_error("HDF5: chunk of " + name + " is shorter than its extent");
            copy_block(data, off, d.chunk, start, count, dst_stride, d.info.elem, sink);
        });
    }

  private:
    static constexpr uint64_t kUndef = ~uint64_t(0);
    struct Filter { int id; };
    struct Dataset {
        VarInfo info;
        int layout = 1;  // 0 compact, 1 contiguous, 2 chunked
        uint64_t addr = kUndef;
        std::vector<uint64_t> chunk;
        std::vector<uint8_t> compact;
        std::vector<Filter> filters;
        bool has_space = false, has_type = false, has_layout = false;
    };
    struct Msg { int type; const uint8_t* p; uint64_t n; };

    const Dataset& set(const std::string& name) const {
        auto it = index_.find(name);
        if (it == index_.end()) throw std::runtime_error("Variable " + name + " not found in file");
        return sets_[it->second];
    }
    uint64_t le(uint64_t off, int n) const {
        const uint8_t* p = f_->at(off, n);
        uint64_t v = 0;
        for (int i = n - 1; i >= 0; --i) v = v << 8 | p[i];
        return v;
    }
    static uint64_t lep(const uint8_t* p, int n) {
        uint64_t v = 0;
        for (int i = n - 1; i >= 0; --i) v = v << 8 | p[i];
        return v;
    }

    // Copy the intersection of block [boff, boff+bshape) with slab [start, start+count) to the sink.
    static void copy_block(const uint8_t* data, const std::vector<uint64_t>& boff, const std::vector<uint64_t>& bshape,
                           const std::vector<uint64_t>& start, const std::vector<uint64_t>& count,
                           const std::vector<uint64_t>& dst_stride, const ElemType& e,
                           const std::function<void(uint64_t, const uint8_t*, uint64_t, const ElemType&)>& sink) {
        const size_t nd = bshape.size();
        std::vector<uint64_t> lo(nd), hi(nd), bstride(nd, 1);
        for (int k = (int)nd - 2; k >= 0; --k) bstride[k] = bstride[k + 1] * bshape[k + 1];
        for (size_t k = 0; k < nd; ++k) {
            lo[k] = std::max(boff[k], start[k]);
            hi[k] = std::min(boff[k] + bshape[k], start[k] + count[k]);
            if (lo[k] >= hi[k]) return;
        }
        std::vector<uint64_t> idx(lo);
        const uint64_t run = hi[nd - 1] - lo[nd - 1];
        for (;;) {
            uint64_t src = 0, dst = 0;
            for (size_t k = 0; k < nd; ++k) {
                src += (idx[k] - boff[k]) * bstride[k];
                dst += (idx[k] - start[k]) * dst_stride[k];
            }
            sink(dst, data + src * e.size, run, e);
            int k = (int)nd - 2;
            for (; k >= 0; --k) {
                if (++idx[k] < hi[k]) break;
                idx[k] = lo[k];
            }
            if (k < 0) break;
        }
    }

    // v1 B-tree of raw-data chunks (node type 1)
    void walk_chunks(const Dataset& d, uint64_t addr,
                     const std::function<void(const std::vector<uint64_t>&, uint64_t, uint32_t, uint32_t)>& fn) const {
        if (addr == kUndef) return;
        const uint8_t* p = f_->at(addr, 8 + 2 * so_);
        if (std::memcmp(p, "TREE", 4) != 0) throw std::runtime_error("HDF5: chunk index of " + d.info.name + " is not a v1 B-tree");
        if (p[4] != 1) throw std::runtime_error("HDF5: unexpected B-tree node type");
        const int level = p[5];
        const int used = (int)lep(p + 6, 2);
        const size_t nd = d.chunk.size();
        const uint64_t key_size = 8 + 8 * (nd + 1);
        uint64_t pos = addr + 8 + 2 * so_;
        for (int i = 0; i < used; ++i) {
            const uint8_t* k = f_->at(pos, key_size + so_);
            const uint32_t size = (uint32_t)lep(k, 4), mask = (uint32_t)lep(k + 4, 4);
            std::vector<uint64_t> off(nd);
            for (size_t j = 0; j < nd; ++j) off[j] = lep(k + 8 + 8 * j, 8);
            const uint64_t child = lep(k + key_size, so_);
            if (level > 0) walk_chunks(d, child, fn);
            else fn(off, child, size, mask);
            pos += key_size + so_;
        }
    }

    // ---- object headers -----------------------------------------------------------------------
    std::vector<Msg> messages(uint64_t addr) const {
        std::vector<Msg> out;
        const uint8_t* p = f_->at(addr, 16);
        if (std::memcmp(p, "OHDR", 4) == 0) {
            if (p[4] != 2) throw std::runtime_error("HDF5: object header version");
            const int flags = p[5];
            uint64_t pos = addr + 6;
            if (flags & 0x20) pos += 16;  // times
            if (flags & 0x10) pos += 4;   // max compact / min dense
            const int szlen = 1 << (flags & 3);
            const uint64_t chunk0 = le(pos, szlen);
            pos += szlen;
            const bool order = flags & 4;
            std::vector<std::pair<uint64_t, uint64_t>> blocks{{pos, chunk0}};
            for (size_t b = 0; b < blocks.size(); ++b) {
                uint64_t q = blocks[b].first, end = blocks[b].first + blocks[b].second;
                while (q + 4 + (order ? 2 : 0) <= end) {
                    const int type = (int)le(q, 1);
                    const uint64_t size = le(q + 1, 2);
                    q += 4 + (order ? 2 : 0);
                    if (q + size > end) break;
                    const uint8_t* body = f_->at(q, size);
                    if (type == 0x10) {  // continuation -> OCHK block: signature, messages, checksum
                        const uint64_t caddr = lep(body, so_), clen = lep(body + so_, sl_);
                        if (std::memcmp(f_->at(caddr, 4), "OCHK", 4) != 0) throw std::runtime_error("HDF5: bad continuation block");
                        blocks.emplace_back(caddr + 4, clen - 8);
                    } else if (type != 0) {
                        out.push_back({type, body, size});
                    }
                    q += size;
                }
            }
            return out;
        }
        // version 1 object header
        if (p[0] != 1) throw std::runtime_error("HDF5: unsupported object header at " + std::to_string(addr));
        const int nmsg = (int)lep(p + 2, 2);
        const uint64_t hsize = lep(p + 8, 4);
        std::vector<std::pair<uint64_t, uint64_t>> blocks{{addr + 16, hsize}};
        int seen = 0;
        for (size_t b = 0; b < blocks.size() && seen < nmsg; ++b) {
            uint64_t q = blocks[b].first, end = blocks[b].first + blocks[b].second;
            while (q + 8 <= end && seen < nmsg) {
                const int type = (int)le(q, 2);
                const uint64_t size = le(q + 2, 2);
                q += 8;
                const uint8_t* body = f_->at(q, size);
                ++seen;
                if (type == 0x10) blocks.emplace_back(lep(body, so_), lep(body + so_, sl_));
                else if (type != 0) out.push_back({type, body, size});
                q += size;
            }
        }
        return out;
    }

    static ElemType parse_datatype(const uint8_t* p, uint64_t n, bool& ok) {
        ElemType e;
        ok = false;
        if (n < 8) return e;
        const int cls = p[0] & 0x0f;
        const int bits0 = p[1];
        e.size = (int)lep(p + 4, 4);
        e.big_endian = bits0 & 1;
        if (cls == 0) {
            e.kind = (bits0 & 8) ? ElemType::kInt : ElemType::kUInt;
            ok = (e.size == 1 || e.size == 2 || e.size == 4 || e.size == 8);
        } else if (cls == 1) {
            e.kind = ElemType::kFloat;
            ok = (e.size == 4 || e.size == 8);
        }
        return e;
    }

    void parse_attribute(const Msg& m, Dataset& d) const {
        const uint8_t* p = m.p;
        const int ver = p[0];
        if (ver < 1 || ver > 3) return;
        const uint64_t name_size = lep(p + 2, 2), dt_size = lep(p + 4, 2), ds_size = lep(p + 6, 2);
        uint64_t q = (ver == 3) ? 9 : 8;
        auto pad = [&](uint64_t x) { return ver == 1 ? ((x + 7) & ~uint64_t(7)) : x; };
        if (q + name_size > m.n) return;
        Attribute a;
        a.name.assign((const char*)p + q, name_size ? name_size - 1 : 0);
        q += pad(name_size);
        const uint8_t* dt = p + q;
        q += pad(dt_size);
        const uint8_t* ds = p + q;
        q += pad(ds_size);
        if (q > m.n) return;
        uint64_t nelem = 1;
        {
            const int sver = ds[0], rank = ds[1];
            const uint64_t off = (sver == 1) ? 8 : 4;
            if (sver == 2 && ds[3] == 2) nelem = 0;  // null dataspace
            for (int k = 0; k < rank; ++k) nelem *= lep(ds + off + (uint64_t)k * sl_, sl_);
        }
        const int cls = dt[0] & 0x0f;
        const uint64_t esize = lep(dt + 4, 4);
        if (cls == 3) {  // fixed-length string
            a.is_text = true;
            const uint64_t len = std::min<uint64_t>(esize * nelem, m.n - q);
            a.text.assign((const char*)p + q, len);
            while (!a.text.empty() && a.text.back() == '\0') a.text.pop_back();
            d.info.atts.push_back(std::move(a));
        } else if (cls == 0 || cls == 1) {
            bool ok;
            const ElemType e = parse_datatype(dt, dt_size, ok);
            if (!ok || q + nelem * e.size > m.n) return;
            for (uint64_t k = 0; k < nelem; ++k) a.numbers.push_back(load_elem<double>(p + q + k * e.size, e));
            d.info.atts.push_back(std::move(a));
        }  // variable-length strings, references (DIMENSION_LIST ...) are skipped
    }

    bool parse_dataset(const std::string& name, uint64_t addr, Dataset& d) const {
        d.info.name = name;
        for (const Msg& m : messages(addr)) {
            const uint8_t* p = m.p;
            if (m.type == 0x01) {  // dataspace
                const int ver = p[0], rank = p[1];
                const uint64_t off = (ver == 1) ? 8 : 4;
                d.info.shape.clear();
                for (int k = 0; k < rank; ++k) d.info.shape.push_back(lep(p + off + (uint64_t)k * sl_, sl_));
                d.info.dim_names.assign(rank, "");
                d.has_space = true;
            } else if (m.type == 0x03) {  // datatype
                bool ok;
                d.info.elem = parse_datatype(p, m.n, ok);
                d.has_type = ok;
            } else if (m.type == 0x08) {  // data layout
                const int ver = p[0];
                if (ver == 3) {
                    d.layout = p[1];
                    if (d.layout == 0) {
                        const uint64_t sz = lep(p + 2, 2);
                        d.compact.assign(p + 4, p + 4 + sz);
                    } else if (d.layout == 1) {
                        d.addr = lep(p + 2, so_);
                    } else if (d.layout == 2) {
                        const int rank = p[2];  // dataset rank + 1
                        d.addr = lep(p + 3, so_);
                        d.chunk.clear();
                        for (int k = 0; k + 1 < rank; ++k) d.chunk.push_back(lep(p + 3 + so_ + 4 * (uint64_t)k, 4));
                    }
                    d.has_layout = true;
                } else if (ver == 1 || ver == 2) {
                    const int rank = p[1];
                    d.layout = p[2];
                    uint64_t q = 8;
                    if (d.layout != 0) { d.addr = lep(p + q, so_); q += so_; }
                    std::vector<uint64_t> dims;
                    for (int k = 0; k < rank; ++k) dims.push_back(lep(p + q + 4 * (uint64_t)k, 4));
                    q += 4 * (uint64_t)rank;
                    if (d.layout == 2) { dims.pop_back(); d.chunk = dims; }
                    if (d.layout == 0) { const uint64_t sz = lep(p + q, 4); d.compact.assign(p + q + 4, p + q + 4 + sz); }
                    d.has_layout = true;
                } else {
                    throw std::runtime_error("HDF5: data layout message version " + std::to_string(ver) + " of " + name +
                                             " is not supported (file written with libver 'latest')");
                }
            } else if (m.type == 0x0B) {  // filter pipeline
                const int ver = p[0], nf = p[1];
                uint64_t q = (ver == 1) ? 8 : 2;
                for (int k = 0; k < nf; ++k) {
                    const int id = (int)lep(p + q, 2);
                    uint64_t name_len = 0;
                    if (ver == 1 || id >= 256) { name_len = lep(p + q + 2, 2); q += 2; }
                    const int ncd = (int)lep(p + q + 4, 2);
                    q += 6;
                    q += (ver == 1) ? ((name_len + 7) & ~uint64_t(7)) : name_len;
                    q += 4 * (uint64_t)ncd;
                    if (ver == 1 && (ncd & 1)) q += 4;
                    d.filters.push_back({id});
                }
            } else if (m.type == 0x0C) {
                parse_attribute(m, d);
            }
        }
        return d.has_space && d.has_type && d.has_layout;
    }

    // Links of a group object: name -> object header address
    void group_links(uint64_t addr, std::vector<std::pair<std::string, uint64_t>>& out) const {
        for (const Msg& m : messages(addr)) {
            const uint8_t* p = m.p;
            if (m.type == 0x11) {  // symbol table: v1 B-tree of SNOD nodes + local heap of names
                const uint64_t btree = lep(p, so_), heap = lep(p + so_, so_);
                const uint8_t* h = f_->at(heap, 8 + 2 * sl_ + so_);
                if (std::memcmp(h, "HEAP", 4) != 0) throw std::runtime_error("HDF5: bad local heap");
                const uint64_t heap_data = lep(h + 8 + 2 * sl_, so_);
                std::function<void(uint64_t)> walk = [&](uint64_t a) {
                    const uint8_t* t = f_->at(a, 8 + 2 * so_);
                    if (std::memcmp(t, "TREE", 4) == 0) {
                        const int used = (int)lep(t + 6, 2);
                        uint64_t pos = a + 8 + 2 * so_ + sl_;  // skip key 0
                        for (int i = 0; i < used; ++i) {
                            walk(le(pos, so_));
                            pos += so_ + sl_;
                        }
                    } else if (std::memcmp(t, "SNOD", 4) == 0) {
                        const int n = (int)lep(t + 6, 2);
                        uint64_t pos = a + 8;
                        for (int i = 0; i < n; ++i) {
                            const uint64_t name_off = le(pos, so_), ohdr = le(pos + so_, so_);
                            const char* nm = (const char*)f_->at(heap_data + name_off, 1);
                            out.emplace_back(std::string(nm), ohdr);
                            pos += 2 * so_ + 4 + 4 + 16;
                        }
                    } else {
                        throw std::runtime_error("HDF5: bad group B-tree node");
                    }
                };
                walk(btree);
            } else if (m.type == 0x06) {  // link message (compact group storage)
                const int flags = p[1];
                uint64_t q = 2;
                int ltype = 0;
                if (flags & 0x08) ltype = p[q++];
                if (flags & 0x04) q += 8;
                if (flags & 0x10) q += 1;
                const int lensz = 1 << (flags & 3);
                const uint64_t nlen = lep(p + q, lensz);
                q += lensz;
                std::string nm((const char*)p + q, nlen);
                q += nlen;
                if (ltype == 0) out.emplace_back(nm, lep(p + q, so_));
            } else if (m.type == 0x02) {  // link info: dense storage (fractal heap) is not read
                const int flags = p[1];
                uint64_t q = 2 + ((flags & 1) ? 8 : 0);
                const uint64_t fheap = lep(p + q, so_);
                if (fheap != kUndef)
                    throw std::runtime_error("HDF5: group with dense link storage (more than 8 objects) is not supported: " + f_->path());
            }
        }
    }

    void parse() {
        const uint8_t* sb = f_->at(0, 16);
        static const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
        if (std::memcmp(sb, sig, 8) != 0) throw std::runtime_error("not an HDF5 file: " + f_->path());
        const int ver = sb[8];
        uint64_t root = kUndef;
        if (ver == 0 || ver == 1) {
            so_ = sb[13];
            sl_ = sb[14];
            uint64_t q = 24 + (ver == 1 ? 4 : 0);
            q += 4 * (uint64_t)so_;            // base, free-space, eof, driver info
            root = le(q + so_, so_);           // root symbol table entry: link name offset, object header address
        } else if (ver == 2 || ver == 3) {
            so_ = sb[9];
            sl_ = sb[10];
            root = le(12 + 3 * (uint64_t)so_, so_);
        } else {
            throw std::runtime_error("HDF5: unsupported superblock version in " + f_->path());
        }
        if (so_ != 8 || sl_ != 8) throw std::runtime_error("HDF5: only 8-byte offsets/lengths are supported");
        std::vector<std::pair<std::string, uint64_t>> links;
        group_links(root, links);
        for (auto& l : links) {
            Dataset d;
            bool ok = false;
            ok = parse_dataset(l.first, l.second, d);
            if (ok) {
                index_[l.first] = sets_.size();
                sets_.push_back(std::move(d));
            }
        }
        // NetCDF-4 names dimensions through dimension-scale datasets: a 1-D dataset whose name equals a
        // coordinate and whose length matches is reported as that dimension's name (best effort, for messages).
        for (auto& d : sets_)
            for (size_t k = 0; k < d.info.shape.size(); ++k)
                for (auto& c : sets_)
                    if (&c != &d && c.info.shape.size() == 1 && c.info.shape[0] == d.info.shape[k] && d.info.dim_names[k].empty() &&
                        c.info.att("CLASS") && c.info.att("CLASS")->text == "DIMENSION_SCALE")
                        d.info.dim_names[k] = c.info.name;
    }

    std::shared_ptr<MappedFile> f_;
    int so_ = 8, sl_ = 8;
    std::vector<Dataset> sets_;
    std::map<std::string, size_t> index_;
};

/// Open a NetCDF file of either container.
inline std::unique_ptr<Reader> open_reader(const std::string& path) {
    auto f = std::make_shared<MappedFile>(path);
    if (f->size() >= 4 && std::memcmp(f->at(0, 3), "CDF", 3) == 0) return std::unique_ptr<Reader>(new ClassicReader(f));
    if (f->size() >= 8 && std::memcmp(f->at(0, 4), "\x89HDF", 4) == 0) return std::unique_ptr<Reader>(new Hdf5Reader(f));
    throw std::runtime_error("Opening file " + path + ": neither NetCDF classic nor NetCDF-4/HDF5");
}

// ---------------------------------------------------------------------------------------------
// NetCDF classic writer (CDF-2)
// ---------------------------------------------------------------------------------------------
class ClassicWriter {
  public:
    explicit ClassicWriter(const std::string& path) : path_(path) {
        fd_ = ::open(path.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644);
        if (fd_ < 0) throw std::runtime_error("NetCDF: cannot create " + path + ": " + std::strerror(errno));
    }
    ~ClassicWriter() { if (fd_ >= 0) ::close(fd_); }
    ClassicWriter(const ClassicWriter&) = delete;
    ClassicWriter& operator=(const ClassicWriter&) = delete;

    int def_dim(const std::string& name, uint64_t len) {
        need_define();
        if (len == 0 || len > 0xFFFFFFFFull) throw std::runtime_error("NetCDF: dimension " + name + " length out of range");
        dims_.push_back({name, len});
        return (int)dims_.size() - 1;
    }
    int def_var(const std::string& name, int type, const std::vector<int>& dimids) {
        need_define();
        V v;
        v.name = name;
        v.type = type;
        v.dimids = dimids;
        vars_.push_back(v);
        return (int)vars_.size() - 1;
    }
    void put_att_text(int varid, const std::string& name, const std::string& text) {
        need_define();
        A a{name, NC_CHAR, std::vector<uint8_t>(text.begin(), text.end()), text.size()};
        (varid < 0 ? gatts_ : vars_.at(varid).atts).push_back(a);
    }
    void put_att_double(int varid, const std::string& name, double value) {
        need_define();
        A a{name, NC_DOUBLE, std::vector<uint8_t>(8), 1};
        uint64_t u;
        std::memcpy(&u, &value, 8);
        for (int i = 0; i < 8; ++i) a.raw[i] = (uint8_t)(u >> (56 - 8 * i));
        (varid < 0 ? gatts_ : vars_.at(varid).atts).push_back(a);
    }
    void enddef() {
        need_define();
        // header size first (begin offsets are 64-bit in CDF-2, so the size does not depend on them)
        std::vector<uint8_t> h = header();
        uint64_t pos = (h.size() + 3) & ~uint64_t(3);
        for (auto& v : vars_) {
            uint64_t n = nc_type_size(v.type);
            for (int d : v.dimids) n *= dims_.at(d).len;
            v.bytes = n;
            v.begin = pos;
            pos += (n + 3) & ~uint64_t(3);
        }
        // CDF-2 allows only the LAST fixed variable to exceed 4 GiB
        for (size_t i = 0; i + 1 < vars_.size(); ++i)
            if (vars_[i].bytes > 0xFFFFFFFCull) throw std::runtime_error("NetCDF CDF-2: only the last variable may exceed 4 GiB");
        h = header();
        total_ = pos;
        if (ftruncate(fd_, (off_t)total_) != 0) throw std::runtime_error("NetCDF: cannot size " + path_);
        pwrite_all(h.data(), h.size(), 0);
        defined_ = true;
    }
    /// Whole variable, host-endian values of the variable's type.
    void put_var(int varid, const void* data) {
        const V& v = vars_.at(varid);
        put_bytes(v, 0, data, v.bytes);
    }
    /// `n` consecutive elements starting at flat element index `first` (C order).
    void put_elems(int varid, uint64_t first, const void* data, uint64_t n) {
        const V& v = vars_.at(varid);
        const uint64_t es = nc_type_size(v.type);
        if ((first + n) * es > v.bytes) throw std::out_of_range("NetCDF: write exceeds variable " + v.name);
        put_bytes(v, first * es, data, n * es);
    }
    /// Writable big-endian view of a variable through a shared mapping (for scattered window writes).
    uint8_t* map_var(int varid) {
        if (!defined_) throw std::runtime_error("NetCDF: enddef() first");
        if (!map_) {
            void* p = mmap(nullptr, total_, PROT_READ | PROT_WRITE, MAP_SHARED, fd_, 0);
            if (p == MAP_FAILED) throw std::runtime_error("NetCDF: cannot map " + path_);
            map_ = (uint8_t*)p;
        }
        return map_ + vars_.at(varid).begin;
    }
    void close() {
        if (map_) { msync(map_, total_, MS_SYNC); munmap(map_, total_); map_ = nullptr; }
        if (fd_ >= 0) { ::close(fd_); fd_ = -1; }
    }
    static void store_be(uint8_t* dst, const void* src, int es, uint64_t n) {
        const uint8_t* s = (const uint8_t*)src;
        if (es == 1) { std::memcpy(dst, s, n); return; }
        if (es == 8) {
            for (uint64_t i = 0; i < n; ++i) {
                uint64_t u;
                std::memcpy(&u, s + 8 * i, 8);
                u = __builtin_bswap64(u);
                std::memcpy(dst + 8 * i, &u, 8);
            }
            return;
        }
        for (uint64_t i = 0; i < n; ++i)
            for (int b = 0; b < es; ++b) dst[i * es + b] = s[i * es + (es - 1 - b)];
    }

  private:
    struct A { std::string name; int type; std::vector<uint8_t> raw; uint64_t nelems; };
    struct D { std::string name; uint64_t len; };
    struct V { std::string name; int type = NC_DOUBLE; std::vector<int> dimids; std::vector<A> atts; uint64_t bytes = 0, begin = 0; };

    void need_define() const { if (defined_) throw std::runtime_error("NetCDF: file is no longer in define mode"); }
    static void p32(std::vector<uint8_t>& h, uint32_t v) { for (int i = 3; i >= 0; --i) h.push_back((uint8_t)(v >> (8 * i))); }
    static void p64(std::vector<uint8_t>& h, uint64_t v) { p32(h, (uint32_t)(v >> 32)); p32(h, (uint32_t)v); }
    static void pname(std::vector<uint8_t>& h, const std::string& s) {
        p32(h, (uint32_t)s.size());
        h.insert(h.end(), s.begin(), s.end());
        while (h.size() & 3) h.push_back(0);
    }
    static void patts(std::vector<uint8_t>& h, const std::vector<A>& atts) {
        if (atts.empty()) { p32(h, 0); p32(h, 0); return; }
        p32(h, 0x0C);
        p32(h, (uint32_t)atts.size());
        for (auto& a : atts) {
            pname(h, a.name);
            p32(h, (uint32_t)a.type);
            p32(h, (uint32_t)a.nelems);
            h.insert(h.end(), a.raw.begin(), a.raw.end());
            while (h.size() & 3) h.push_back(0);
        }
    }
    std::vector<uint8_t> header() const {
        std::vector<uint8_t> h{'C', 'D', 'F', 2};
        p32(h, 0);  // numrecs: no record variables
        if (dims_.empty()) { p32(h, 0); p32(h, 0); }
        else {
            p32(h, 0x0A);
            p32(h, (uint32_t)dims_.size());
            for (auto& d : dims_) { pname(h, d.name); p32(h, (uint32_t)d.len); }
        }
        patts(h, gatts_);
        if (vars_.empty()) { p32(h, 0); p32(h, 0); }
        else {
            p32(h, 0x0B);
            p32(h, (uint32_t)vars_.size());
            for (auto& v : vars_) {
                pname(h, v.name);
                p32(h, (uint32_t)v.dimids.size());
                for (int d : v.dimids) p32(h, (uint32_t)d);
                patts(h, v.atts);
                p32(h, (uint32_t)v.type);
                const uint64_t padded = (v.bytes + 3) & ~uint64_t(3);
                p32(h, padded > 0xFFFFFFFCull ? 0xFFFFFFFFu : (uint32_t)padded);
                p64(h, v.begin);
            }
        }
        return h;
    }
    void pwrite_all(const void* p, uint64_t n, uint64_t off) {
        const uint8_t* b = (const uint8_t*)p;
        while (n) {
            const ssize_t w = ::pwrite(fd_, b, n, (off_t)off);
            if (w <= 0) throw std::runtime_error("NetCDF: write to " + path_ + " failed: " + std::strerror(errno));
            b += w; off += (uint64_t)w; n -= (uint64_t)w;
        }
    }
    void put_bytes(const V& v, uint64_t byte_off, const void* data, uint64_t nbytes) {
        if (!defined_) throw std::runtime_error("NetCDF: enddef() first");
        const int es = nc_type_size(v.type);
        const uint64_t block = 1 << 20;
        std::vector<uint8_t> tmp(std::min<uint64_t>(nbytes, block * es));
        const uint8_t* s = (const uint8_t*)data;
        uint64_t done = 0;
        while (done < nbytes) {
            const uint64_t n = std::min<uint64_t>(nbytes - done, tmp.size());
            store_be(tmp.data(), s + done, es, n / es);
            pwrite_all(tmp.data(), n, v.begin + byte_off + done);
            done += n;
        }
    }

    std::string path_;
    int fd_ = -1;
    bool defined_ = false;
    std::vector<D> dims_;
    std::vector<A> gatts_;
    std::vector<V> vars_;
    uint64_t total_ = 0;
    uint8_t* map_ = nullptr;
};

}  // namespace hlmnc

// ---------------------------------------------------------------------------------------------
// The reference's names
// ---------------------------------------------------------------------------------------------

/// I_O/forcing_loader.hpp:34-89 — (time, lat, lon) float variable read by time chunks.
class NetCDFLoader {
  public:
    NetCDFLoader(const std::string& filename, const std::string& varName) : fileName(filename), varName(varName) {
        reader_ = hlmnc::open_reader(filename);  // throws "Opening file ..." like forcing_loader.cpp:80-81
        if (!reader_->has_variable(varName)) throw std::runtime_error("Variable " + varName + " not found in file");
        const hlmnc::VarInfo& v = reader_->inquire(varName);
        if (v.shape.size() != 3)
            throw std::runtime_error("Expected 3D variable (time, lat, lon), got " + std::to_string(v.shape.size()) + "D");
        timeSize = v.shape[0];
        latSize = v.shape[1];
        lonSize = v.shape[2];
        // stdio, not iostream: a plug-in loaded into a process with an older libstdc++ must not depend
        // on that library having initialised std::cout
        if (verbose) std::printf("Dataset dimensions: %zu x %zu x %zu\n", timeSize, latSize, lonSize);
    }
    NetCDFLoader(NetCDFLoader&&) = default;
    NetCDFLoader& operator=(NetCDFLoader&&) = default;

    /// forcing_loader.cpp:165-196; same argument checks and exception types.
    std::unique_ptr<float[]> loadTimeChunk(size_t startTime, size_t numTimeSteps) {
        if (numTimeSteps == 0) throw std::invalid_argument("Size of time chunk must be greater than zero");
        if (startTime >= timeSize) throw std::out_of_range("Start time index out of range");
        if (startTime + numTimeSteps > timeSize) throw std::out_of_range("Requested time steps exceed available data");
        std::unique_ptr<float[]> data(new float[numTimeSteps * latSize * lonSize]);
        reader_->read_into<float>(varName, {startTime, 0, 0}, {numTimeSteps, latSize, lonSize}, data.get());
        if (verbose)
            std::printf("Loaded time chunk: steps %zu to %zu (%zu time steps)\n", startTime, startTime + numTimeSteps - 1, numTimeSteps);
        return data;
    }
    /// forcing_loader.cpp:199-212
    float getValueFromChunk(const std::unique_ptr<float[]>& chunkData, size_t relativeTimeIndex, size_t latIndex,
                            size_t lonIndex, size_t chunkTimeSize, size_t latSz, size_t lonSz) {
        if (relativeTimeIndex >= chunkTimeSize || latIndex >= latSz || lonIndex >= lonSz)
            throw std::out_of_range("Chunk indices out of range");
        return chunkData[relativeTimeIndex * (latSz * lonSz) + latIndex * lonSz + lonIndex];
    }
    bool isDataLoaded() const { return reader_ && timeSize > 0 && latSize > 0 && lonSize > 0; }
    size_t getTimeSize() const { return timeSize; }
    size_t getLatSize() const { return latSize; }
    size_t getLonSize() const { return lonSize; }
    std::string getVariableName() const { return varName; }
    std::string getFileName() const { return fileName; }
    const hlmnc::Reader& reader() const { return *reader_; }

    /// Sample spacing in hours from the file's time coordinate ("<unit> since ..."), or 0 when the file
    /// does not say (the reference hard-codes 1 h / 24 h, main.cpp:520-523).
    double timeStepHours() const {
        const hlmnc::VarInfo& v = reader_->inquire(varName);
        std::string tname = v.dim_names.empty() ? "" : v.dim_names[0];
        if (tname.empty() || !reader_->has_variable(tname)) tname = reader_->has_variable("time") ? "time" : (reader_->has_variable("valid_time") ? "valid_time" : "");
        if (tname.empty()) return 0.0;
        const hlmnc::VarInfo& t = reader_->inquire(tname);
        if (t.shape.size() != 1 || t.shape[0] < 2) return 0.0;
        const auto* units = t.att("units");
        if (!units || !units->is_text) return 0.0;
        double per_hour = 0.0;
        const std::string& u = units->text;
        if (u.compare(0, 6, "second") == 0) per_hour = 3600.0;
        else if (u.compare(0, 6, "minute") == 0) per_hour = 60.0;
        else if (u.compare(0, 4, "hour") == 0) per_hour = 1.0;
        else if (u.compare(0, 3, "day") == 0) per_hour = 1.0 / 24.0;
        if (per_hour == 0.0) return 0.0;
        const std::vector<double> tv = reader_->read_as<double>(tname, {0}, {2});
        return (tv[1] - tv[0]) / per_hour;
    }
    bool verbose = true;

  private:
    std::unique_ptr<hlmnc::Reader> reader_;
    size_t timeSize = 0, latSize = 0, lonSize = 0;
    std::string fileName, varName;
};

/// Container of an output file: NetCDF-4 (HDF5; what the reference writes, I_O/output_series.cpp:31,88) or the
/// classic 64-bit-offset format (no compression; every NetCDF reader, SciPy's included, opens it).
enum class NcFormat { kNetcdf4, kClassic };

namespace hlmnc_detail {
// the reference's dimensions, coordinate variables and attributes (output_series.cpp:31-52,88-105), on either writer
template <class W> struct SeriesVars { int vs, vt, vv, vo; };
template <class W>
SeriesVars<W> define_series(W& w, uint64_t ns, uint64_t nq, uint64_t nv, bool with_time, int out_type) {
    SeriesVars<W> r{};
    const int ds = w.def_dim("system", ns);
    const int dt = with_time ? w.def_dim("time", nq) : -1;
    const int dv = w.def_dim("variable", nv);
    r.vs = w.def_var("system", hlmnc::NC_INT, {ds});
    r.vt = with_time ? w.def_var("time", hlmnc::NC_DOUBLE, {dt}) : -1;
    r.vv = w.def_var("variable", hlmnc::NC_INT, {dv});
    w.put_att_text(r.vs, "long_name", "LinkID");
    if (with_time) {
        w.put_att_text(r.vt, "long_name", "Time");
        w.put_att_text(r.vt, "units", "minutes since start of simulation");
    }
    w.put_att_text(r.vv, "long_name", "state variable");
    w.put_att_text(r.vv, "units", "various units");
    r.vo = with_time ? w.def_var("outputs", out_type, {ds, dt, dv}) : w.def_var("outputs", out_type, {ds, dv});
    return r;
}
}  // namespace hlmnc_detail

/// I_O/output_series.cpp:18-71 — outputs(system,time,variable) with coordinate variables and the
/// reference's attributes; NetCDF-4 with shuffle + deflate(compression_level) on `outputs` as the reference
/// (compression_level 0: no filter), or the classic container.  Errors are printed and swallowed like the
/// reference's NC_CHECK.
inline void write_dense_netcdf(const std::string& filename, const double* h_dense, const double* time_vals,
                               const int* linkid_vals, const int* state_vals, int num_queries, int num_systems, int N_EQ,
                               int compression_level = 4, NcFormat format = NcFormat::kNetcdf4) {
    try {
        if (format == NcFormat::kNetcdf4) {
            hlmnc::Nc4Writer w(filename);
            const auto v = hlmnc_detail::define_series(w, num_systems, num_queries, N_EQ, true, hlmnc::NC_DOUBLE);
            w.def_var_deflate(v.vo, true, compression_level);
            w.put_var(v.vs, linkid_vals);
            w.put_var(v.vt, time_vals);
            w.put_var(v.vv, state_vals);
            w.put_var(v.vo, h_dense);
            w.close();
            return;
        }
        hlmnc::ClassicWriter w(filename);
        const auto v = hlmnc_detail::define_series(w, num_systems, num_queries, N_EQ, true, hlmnc::NC_DOUBLE);
        w.enddef();
        w.put_var(v.vs, linkid_vals);
        w.put_var(v.vt, time_vals);
        w.put_var(v.vv, state_vals);
        w.put_var(v.vo, h_dense);
        w.close();
    } catch (const std::exception& e) {
        std::fprintf(stderr, "NetCDF error: %s\n", e.what());
    }
}

/// I_O/output_series.cpp:76-123 — outputs(system,variable).
inline void write_final_netcdf(const std::string& filename, const double* h_y_final, const int* linkid_vals,
                               const int* state_vals, int num_systems, int N_EQ, int compression_level = 4,
                               NcFormat format = NcFormat::kNetcdf4) {
    try {
        if (format == NcFormat::kNetcdf4) {
            hlmnc::Nc4Writer w(filename);
            const auto v = hlmnc_detail::define_series(w, num_systems, 0, N_EQ, false, hlmnc::NC_DOUBLE);
            w.def_var_deflate(v.vo, true, compression_level);
            w.put_var(v.vs, linkid_vals);
            w.put_var(v.vv, state_vals);
            w.put_var(v.vo, h_y_final);
            w.close();
            return;
        }
        hlmnc::ClassicWriter w(filename);
        const auto v = hlmnc_detail::define_series(w, num_systems, 0, N_EQ, false, hlmnc::NC_DOUBLE);
        w.enddef();
        w.put_var(v.vs, linkid_vals);
        w.put_var(v.vv, state_vals);
        w.put_var(v.vo, h_y_final);
        w.close();
    } catch (const std::exception& e) {
        std::fprintf(stderr, "NetCDF error: %s\n", e.what());
    }
}

/// outputs(system,time,variable) written window by window: the solver hands over blocks of consecutive query
/// ranges (hlm_solve_fetch_window_packed) and only the selected states are kept (config.yaml output.states).
/// Two source layouts: records of all n_eq states, of which the writer picks its own (the default), or — after
/// set_packed_source() — records the DEVICE already cut down to the selected states in ascending state order
/// (hlm_set_output_states), as double or float (hlm_set_output_precision; the file variable is then NC_FLOAT).
/// Two containers:
///   * classic: rows are scattered into a shared mapping of the file, so a window costs one pass over its bytes;
///   * NetCDF-4 (the reference's, default): `outputs` is chunked (systems x queries x all variables) with shuffle +
///     deflate like the reference's; windows fill a staging slab of one chunk-row of queries, and a complete slab is
///     compressed chunk by chunk on the host's threads and appended — the file never exists uncompressed.  A series
///     that fits one 4 MiB chunk is a single chunk, as in the reference's files.
class DenseSeriesWriter {
  public:
    DenseSeriesWriter(const std::string& filename, const std::vector<double>& time_vals, const std::vector<int>& linkid_vals,
                      const std::vector<int>& state_vals, int n_eq, bool store_float = false, NcFormat format = NcFormat::kNetcdf4,
                      int compression_level = 4, uint64_t staging_bytes = 256ULL << 20)
        : nq_(time_vals.size()), ns_(linkid_vals.size()), n_eq_(n_eq), states_(state_vals), f32_(store_float) {
        for (int s : states_)
            if (s < 0 || s >= n_eq) throw std::runtime_error("output state index out of range");
        src_cols_ = states_;
        src_stride_ = (uint64_t)n_eq;
        const int out_type = f32_ ? hlmnc::NC_FLOAT : hlmnc::NC_DOUBLE;
        const uint64_t nv = states_.size(), es = f32_ ? 4 : 8;
        if (format == NcFormat::kClassic) {
            w_.reset(new hlmnc::ClassicWriter(filename));
            const auto v = hlmnc_detail::define_series(*w_, ns_, nq_, nv, true, out_type);
            vo_ = v.vo;
            w_->enddef();
            w_->put_var(v.vs, linkid_vals.data());
            w_->put_var(v.vt, time_vals.data());
            w_->put_var(v.vv, states_.data());
            out_ = w_->map_var(vo_);
            return;
        }
        n4_.reset(new hlmnc::Nc4Writer(filename));
        const auto v = hlmnc_detail::define_series(*n4_, ns_, nq_, nv, true, out_type);
        vo_ = v.vo;
        // chunk = (cs systems, cq queries, all variables): cq queries of every system must fit the staging slab, a chunk
        // aims at netcdf-c's 4 MiB default; one chunk when the whole series fits
        const uint64_t per_q = std::max<uint64_t>(ns_ * nv * es, 1);
        cq_ = std::max<uint64_t>(1, std::min<uint64_t>(std::max<uint64_t>(nq_, 1), staging_bytes / per_q));
        cs_ = std::max<uint64_t>(1, std::min<uint64_t>(std::max<uint64_t>(ns_, 1), hlmnc::Nc4Writer::kDefaultChunkBytes / std::max<uint64_t>(cq_ * nv * es, 1)));
        n4_->def_var_deflate(vo_, true, compression_level, {cs_, cq_, std::max<uint64_t>(nv, 1)});
        n4_->put_var(v.vs, linkid_vals.data());
        n4_->put_var(v.vt, time_vals.data());
        n4_->put_var(v.vv, states_.data());
        slab_.resize(ns_ * cq_ * nv * es);
        hlmnc::Nc4Writer::fill_with_default(out_type, slab_.data(), ns_ * cq_ * nv);
    }
    /// The source records hold only the distinct selected states, ascending (what the device writes under
    /// hlm_set_output_states(output_mask())); values are float when the writer stores float.
    void set_packed_source() {
        std::vector<int> sorted = states_;
        std::sort(sorted.begin(), sorted.end());
        sorted.erase(std::unique(sorted.begin(), sorted.end()), sorted.end());
        for (size_t v = 0; v < states_.size(); ++v)
            src_cols_[v] = (int)(std::lower_bound(sorted.begin(), sorted.end(), states_[v]) - sorted.begin());
        src_stride_ = sorted.size();
        packed_ = true;
    }
    /// bit i set = state i is written: the argument of hlm_set_output_states
    unsigned int output_mask() const {
        unsigned int m = 0;
        for (int s : states_) m |= 1u << s;
        return m;
    }
    /// win = [ns][q_hi - q_lo][source record] (row pitch given in queries) holding queries [q_lo, q_hi);
    /// values are double, or float for a packed float source.  Windows arrive in ascending, gap-free order.
    void write_window(const void* win, uint64_t q_lo, uint64_t q_hi, uint64_t pitch_q) {
        if (q_hi > nq_ || q_lo > q_hi) throw std::out_of_range("dense window outside the query range");
        const uint64_t nv = states_.size();
        const bool src_f32 = packed_ && f32_;
        const int es = f32_ ? 4 : 8;
        auto put = [&](uint8_t* dst, uint64_t rec, bool big_endian) {
            for (uint64_t v = 0; v < nv; ++v) {
                if (src_f32) {
                    const float* p = static_cast<const float*>(win) + rec + src_cols_[v];
                    if (big_endian) hlmnc::ClassicWriter::store_be(dst + 4 * v, p, 4, 1);
                    else std::memcpy(dst + 4 * v, p, 4);
                } else if (f32_) {
                    const float x = (float)static_cast<const double*>(win)[rec + src_cols_[v]];
                    if (big_endian) hlmnc::ClassicWriter::store_be(dst + 4 * v, &x, 4, 1);
                    else std::memcpy(dst + 4 * v, &x, 4);
                } else {
                    const double* p = static_cast<const double*>(win) + rec + src_cols_[v];
                    if (big_endian) hlmnc::ClassicWriter::store_be(dst + 8 * v, p, 8, 1);
                    else std::memcpy(dst + 8 * v, p, 8);
                }
            }
        };
        if (w_) {
            for (uint64_t s = 0; s < ns_; ++s)
                for (uint64_t q = q_lo; q < q_hi; ++q)
                    put(out_ + ((s * nq_ + q) * nv) * es, (s * pitch_q + (q - q_lo)) * src_stride_, true);
            return;
        }
        if (q_lo != q_next_) throw std::runtime_error("NetCDF-4 dense writer: windows must arrive in order, without gaps");
        for (uint64_t qa = q_lo; qa < q_hi;) {  // the part of the window inside the current chunk-row of queries
            const uint64_t row = qa / cq_, qb = std::min(q_hi, (row + 1) * cq_);
            for (uint64_t s = 0; s < ns_; ++s)
                for (uint64_t q = qa; q < qb; ++q)
                    put(slab_.data() + ((s * cq_ + (q - row * cq_)) * nv) * es, (s * pitch_q + (q - q_lo)) * src_stride_, false);
            qa = qb;
            q_next_ = qb;
            if (qb == (row + 1) * cq_ || qb == nq_) flush_slab(row);
        }
    }
    void close() {
        if (w_) w_->close();
        if (n4_) n4_->close();
    }

  private:
    // compress the chunks of one chunk-row (all systems x cq_ queries) on the host's threads, append them in order
    void flush_slab(uint64_t row) {
        const uint64_t nv = states_.size(), es = f32_ ? 4 : 8;
        const uint64_t n_chunks = (ns_ + cs_ - 1) / cs_, chunk_bytes = cs_ * cq_ * nv * es;
        const int out_type = f32_ ? hlmnc::NC_FLOAT : hlmnc::NC_DOUBLE;
        std::vector<std::vector<uint8_t>> enc(n_chunks);
        const unsigned n_thr = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>({(uint64_t)std::max(1u, std::thread::hardware_concurrency()), n_chunks, (uint64_t)32}));
        std::vector<std::thread> pool;
        std::string err;
        for (unsigned t = 0; t < n_thr; ++t)
            pool.emplace_back([&, t] {
                try {
                    std::vector<uint8_t> padded;
                    for (uint64_t c = t; c < n_chunks; c += n_thr) {
                        const uint64_t s0 = c * cs_, n_s = std::min(cs_, ns_ - s0);
                        const uint8_t* src = slab_.data() + s0 * cq_ * nv * es;
                        if (n_s < cs_) {  // the last chunk overhangs the system dimension: pad with the fill value
                            padded.resize(chunk_bytes);
                            hlmnc::Nc4Writer::fill_with_default(out_type, padded.data(), cs_ * cq_ * nv);
                            std::memcpy(padded.data(), src, n_s * cq_ * nv * es);
                            src = padded.data();
                        }
                        n4_->encode_chunk(vo_, src, enc[c]);
                    }
                } catch (const std::exception& e) {
                    err = e.what();
                }
            });
        for (auto& th : pool) th.join();
        if (!err.empty()) throw std::runtime_error(err);
        for (uint64_t c = 0; c < n_chunks; ++c) n4_->put_encoded_chunk(vo_, {c, row, 0}, enc[c]);
        hlmnc::Nc4Writer::fill_with_default(out_type, slab_.data(), ns_ * cq_ * nv);  // queries past nq_ in the last row stay fill
    }

    std::unique_ptr<hlmnc::ClassicWriter> w_;
    std::unique_ptr<hlmnc::Nc4Writer> n4_;
    uint64_t nq_, ns_;
    int n_eq_, vo_ = -1;
    std::vector<int> states_, src_cols_;
    uint64_t src_stride_ = 0;
    bool f32_ = false, packed_ = false;
    uint8_t* out_ = nullptr;
    // NetCDF-4: chunk shape (cs_ systems x cq_ queries x all variables), staging slab [ns][cq_][nv], next query expected
    uint64_t cs_ = 1, cq_ = 1, q_next_ = 0;
    std::vector<uint8_t> slab_;
};
