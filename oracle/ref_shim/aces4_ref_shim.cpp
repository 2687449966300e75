// oracle/ref_shim/aces4_ref_shim.cpp -- TEST INFRASTRUCTURE.  Never linked into, imported by or shipped with the
// product (libsipgpu.so); only tests/ load the library this file is part of.
//
// A plain C ABI over the parts of the REFERENCE's OWN C++ that sit on the hot path and compile with g++ alone
// (no MPI, no BLAS, no Fortran, no cmake).  oracle/Makefile compiles those reference sources where they lie
// under /root/reference/src/sip -- nothing is copied into this repository -- together with this file into
// oracle/_ref/libaces4_ref.so.  The tests use it to pin the C restatement (oracle/tensor_dil_oracle.c) and the
// product's host logic against reference code instead of against a reading of it:
//
//   sip::Block::fill / scale / scale_and_copy / copy_data_ / increment_elements / accumulate_data
//                                                        src/sip/dynamic_data/block.cpp:132-268
//   sip::Block::transpose_copy (permutation-vector conversion) / extract_slice / insert_slice
//                                                        block.cpp:216-255, 272-323
//   sip::BlockShape::num_elems, sip::BlockId ordering     block_shape.cpp, block_id.cpp
//   sip::DistributedBlockConsistency::update_and_check_consistency
//                                                        distributed_block_consistency.cpp:25-175
//   sip::ArrayTableEntry::init_calculated_values / block_number / num2id
//                                                        src/sip/static_data/array_table.cpp:50-97
//   setup::SetupReader + setup::BinaryInputFile / BinaryOutputFile (.dat stream format)
//                                                        src/sip/setup/setup_reader.cpp:329-612, io_utils.cpp
//
// What is NOT reference code in that library: the three Fortran kernels block.cpp calls
// (tensor_block_copy__/slice__/insert__, tensor_dil_omp.F90) -- there is no Fortran compiler in this image, so
// they are forwarded to the oracle's C restatement; tests through this path therefore pin the C++ layer above the
// Fortran (argument conversion, shapes, offsets), not the Fortran loop nests -- and the link-closure stubs at the
// bottom of this file (diagnostic name lookups of SipTables, current_line, Interpreter::global_interpreter).
#include <cstring>
#include <cstdio>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
#include <bitset>
#include <iostream>
#include <fstream>
#include <set>
#include <stack>
#include <list>
#include <algorithm>

// The static tables can only be populated from a .siox file in the reference (private init paths, friend readers), and
// SetupReader keeps its segment table private with an accessor that is declared but never defined; the shim reads /
// fills those few fields directly.  Access control does not change object layout.
#define private public
#include "config.h"
#include "sip.h"
#include "index_table.h"
#include "array_table.h"
#include "block.h"
#include "block_id.h"
#include "block_shape.h"
#include "memory_tracker.h"
#include "job_control.h"
#include "distributed_block_consistency.h"
#include "sip_mpi_constants.h"
#include "io_utils.h"
#include "setup_reader.h"
#include "sip_tables.h"
#include "interpreter.h"
#include "sip_interface.h"
#undef private
#ifdef HAVE_CUDA
#include "gpu_super_instructions.h"   // level-1 build: ref_shim/level1/gpu_super_instructions.h -> sipgpu.h
#endif


// ---- the three Fortran kernels block.cpp calls ----
// default build (libaces4_ref.so): forwarded to the oracle's C restatement (see header);
// -DACES4_REF_LINK_SIPGPU (libaces4_ref_on_sipgpu.so): left undefined here and resolved by the PRODUCT, libsipgpu.so,
// which exports them with the reference's names -- the reference's Block methods then run, unmodified, on the CUDA library
// (INTEGRATION.md level 0), which is what tests/test_gpu_ref_block_on_sipgpu.py checks.
#ifndef ACES4_REF_LINK_SIPGPU
extern "C" {
void oracle_tensor_block_copy__(const int* nthreads, const int* rank, const int* ext, const int* transp,
                                const double* in, double* out, int* ierr);
void oracle_tensor_block_slice__(const int* nthreads, const int* rank, const double* t, const int* t_ext, double* s,
                                 const int* s_ext, const int* beg0, int* ierr);
void oracle_tensor_block_insert__(const int* nthreads, const int* rank, double* t, const int* t_ext, const double* s,
                                  const int* s_ext, const int* beg0, int* ierr);

void tensor_block_copy__(int& nthreads, int& rank, int* ext, int* transp, double* in, double* out, int& ierr) {
    oracle_tensor_block_copy__(&nthreads, &rank, ext, transp, in, out, &ierr);
}
void tensor_block_slice__(int& nthreads, int& rank, double* t, int* t_ext, double* s, int* s_ext, int* beg0, int& ierr) {
    oracle_tensor_block_slice__(&nthreads, &rank, t, t_ext, s, s_ext, beg0, &ierr);
}
void tensor_block_insert__(int& nthreads, int& rank, double* t, int* t_ext, double* s, int* s_ext, int* beg0, int& ierr) {
    oracle_tensor_block_insert__(&nthreads, &rank, t, t_ext, s, s_ext, beg0, &ierr);
}
}
#endif

namespace {

std::string g_last_error;

void ensure_globals() {
    if (sip::JobControl::global == NULL) sip::JobControl::set_global_job_control(new sip::JobControl(std::string("aces4_b200_oracle_ref")));
    if (sip::MemoryTracker::global == NULL) sip::MemoryTracker::set_global_memory_tracker(new sip::MemoryTracker());
}

sip::BlockShape make_shape(int rank, const int* ext) {
    sip::segment_size_array_t s;
    for (int i = 0; i < MAX_RANK; ++i) s[i] = i < rank ? ext[i] : 1;
    return sip::BlockShape(s, rank);
}

// a reference Block holding a copy of caller data (the Block owns and frees its array)
sip::Block* make_block(int rank, const int* ext, const double* data) {
    sip::Block* b = new sip::Block(make_shape(rank, ext));
    if (data) std::memcpy(b->get_data(), data, sizeof(double) * (size_t)b->size());
    return b;
}

template <class F> int guarded(F f) {
    try {
        ensure_globals();
        f();
        return 0;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return 1;
    } catch (...) {
        g_last_error = "unknown exception";
        return 1;
    }
}

}  // namespace

extern "C" {

const char* aces4ref_last_error() { return g_last_error.c_str(); }

int aces4ref_max_rank() { return MAX_RANK; }

long long aces4ref_shape_num_elems(int rank, const int* ext) {
    long long n = -1;
    guarded([&] { n = make_shape(rank, ext).num_elems(); });
    return n;
}

// op: 0 fill(x)  1 scale(x)  2 scale_and_copy(src, x)  3 copy_data_(src)  4 increment_elements(x)  5 accumulate_data(src)
int aces4ref_block_op(int op, int rank, const int* ext, double* d, const double* src, double x) {
    return guarded([&] {
        sip::Block* db = make_block(rank, ext, d);
        sip::Block* sb = src ? make_block(rank, ext, src) : NULL;
        switch (op) {
        case 0: db->fill(x); break;
        case 1: db->scale(x); break;
        case 2: db->scale_and_copy(sb, x); break;
        case 3: db->copy_data_(sb); break;
        case 4: db->increment_elements(x); break;
        case 5: db->accumulate_data(sb); break;
        default: delete db; delete sb; throw std::invalid_argument("aces4ref_block_op: bad op");
        }
        std::memcpy(d, db->get_data(), sizeof(double) * (size_t)db->size());
        delete db;
        delete sb;
    });
}

// Block::transpose_copy: permute[i] (0-based) is handed over exactly as Interpreter::permute_rhs_to_lhs builds it
int aces4ref_block_transpose_copy(int rank, const int* src_ext, const int* dst_ext, const int* permute, const double* src,
                                  double* dst) {
    return guarded([&] {
        sip::Block* sb = make_block(rank, src_ext, src);
        sip::Block* db = make_block(rank, dst_ext, NULL);
        sip::Block::permute_t p;
        for (int i = 0; i < MAX_RANK; ++i) p[i] = permute[i];
        db->transpose_copy(sb, rank, p);
        std::memcpy(dst, db->get_data(), sizeof(double) * (size_t)db->size());
        delete sb;
        delete db;
    });
}

int aces4ref_block_extract_slice(int rank, const int* t_ext, const double* t, const int* s_ext, const int* offsets, double* s) {
    return guarded([&] {
        sip::Block* tb = make_block(rank, t_ext, t);
        sip::Block* sb = make_block(rank, s_ext, NULL);
        sip::offset_array_t off;
        for (int i = 0; i < MAX_RANK; ++i) off[i] = i < rank ? offsets[i] : 0;
        tb->extract_slice(rank, off, sb);
        std::memcpy(s, sb->get_data(), sizeof(double) * (size_t)sb->size());
        delete tb;
        delete sb;
    });
}

int aces4ref_block_insert_slice(int rank, const int* t_ext, double* t, const int* s_ext, const int* offsets, const double* s) {
    return guarded([&] {
        sip::Block* tb = make_block(rank, t_ext, t);
        sip::Block* sb = make_block(rank, s_ext, s);
        sip::offset_array_t off;
        for (int i = 0; i < MAX_RANK; ++i) off[i] = i < rank ? offsets[i] : 0;
        tb->insert_slice(rank, off, sb);
        std::memcpy(t, tb->get_data(), sizeof(double) * (size_t)tb->size());
        delete tb;
        delete sb;
    });
}

// -1: a < b, 0: equal, 1: a > b under the reference's BlockId::operator< / operator== (the block-map key order)
int aces4ref_block_id_compare(int array_a, const int* idx_a, int array_b, const int* idx_b) {
    int r = -2;
    guarded([&] {
        sip::index_value_array_t a, b;
        for (int i = 0; i < MAX_RANK; ++i) { a[i] = idx_a[i]; b[i] = idx_b[i]; }
        sip::BlockId ia(array_a, a), ib(array_b, b);
        r = ia == ib ? 0 : (ia < ib ? -1 : 1);
    });
    return r;
}

// One server block, a sequence of accesses: ops[i] in {0 GET, 1 PUT, 2 PUT_ACCUMULATE} by workers[i] in barrier section
// sections[i].  Returns the index of the first access update_and_check_consistency rejects, -1 if all are accepted.
long long aces4ref_block_consistency(long long n, const int* ops, const int* workers, const int* sections) {
    long long bad = -1;
    int rc = guarded([&] {
        sip::DistributedBlockConsistency c;
        for (long long i = 0; i < n; ++i) {
            sip::SIPMPIConstants::MessageType_t m = ops[i] == 0   ? sip::SIPMPIConstants::GET
                                                    : ops[i] == 1 ? sip::SIPMPIConstants::PUT
                                                                  : sip::SIPMPIConstants::PUT_ACCUMULATE;
            if (!c.update_and_check_consistency(m, workers[i], sections[i])) { bad = i; return; }
        }
    });
    return rc ? -2 : bad;
}

// A rank-`rank` distributed array whose i-th index has nseg[i] segments numbered from lower[i] (extents are irrelevant to
// the numbering; every segment gets extent 1 + its position).  Runs ArrayTableEntry::init_calculated_values on a real
// IndexTable and returns block_number(idx); num2id is returned in idx_back[rank].
long long aces4ref_block_number(int rank, const int* nseg, const int* lower, const int* idx, int* idx_back) {
    long long num = -1;
    guarded([&] {
        sip::IndexTable table;
        std::vector<sip::SegmentDescriptor*> owned;
        sip::index_selector_t sel;
        for (int i = 0; i < MAX_RANK; ++i) sel[i] = sip::unused_index_slot;
        for (int i = 0; i < rank; ++i) {
            // the descriptor is addressed by segment VALUE (1-based); pad so that values up to lower+nseg-1 exist
            std::vector<int> extents;
            for (int s = 0; s < lower[i] - 1 + nseg[i]; ++s) extents.push_back(1 + s);
            sip::IndexTableEntry e;
            e.name_ = "i" + std::to_string(i);
            e.index_type_ = sip::moaindex;
            e.lower_seg_ = lower[i];
            e.num_segments_ = nseg[i];
            e.segment_descriptor_ptr_ = new sip::NonuniformSegmentDescriptor(extents);
            owned.push_back(e.segment_descriptor_ptr_);
            table.entries_.push_back(e);
            sel[i] = i;
        }
        sip::ArrayTableEntry entry("A", rank, sip::distributed_array_t, sel, -1);
        entry.init_calculated_values(table);
        sip::index_value_array_t v;
        for (int i = 0; i < MAX_RANK; ++i) v[i] = i < rank ? idx[i] : sip::unused_index_value;
        sip::BlockId id(0, v);
        num = (long long)entry.block_number(id);
        sip::BlockId back = entry.num2id(0, (size_t)num);
        for (int i = 0; i < rank; ++i) idx_back[i] = back.index_values(i);
        for (size_t k = 0; k < owned.size(); ++k) delete owned[k];
        for (size_t k = 0; k < table.entries_.size(); ++k) table.entries_[k].segment_descriptor_ptr_ = NULL;
    });
    return num;
}

#ifdef HAVE_CUDA
// ---- INTEGRATION.md level 1: the device half of sip::Block (block.cpp:377-429, compiled with HAVE_CUDA against the
// replacement gpu_super_instructions.h of ref_shim/level1) on the product's `_gpu_*` entry points.
// new_gpu_block (zero-filled by _gpu_allocate) -> gpu_fill(fill) -> gpu_scale(scale) -> a second block gpu_copy_data's it;
// both are read back with the product's _gpu_device_to_host; returns 0, or 2 when no device block could be allocated. ----
int aces4ref_gpu_block_roundtrip(int rank, const int* ext, double fill, double scale, double* fresh, double* filled_scaled,
                                 double* copied) {
    int rc = 0;
    int g = guarded([&] {
        sip::Block* a = sip::Block::new_gpu_block(make_shape(rank, ext));
        sip::Block* b = sip::Block::new_gpu_block(make_shape(rank, ext));
        if (!a->get_gpu_data() || !b->get_gpu_data() || !a->is_on_gpu() || a->is_on_host()) {
            rc = 2;
        } else {
            const int n = a->size();
            if (_gpu_device_to_host(fresh, a->get_gpu_data(), n)) rc = 3;
            a->gpu_fill(fill);
            a->gpu_scale(scale);
            b->gpu_copy_data(a);
            if (_gpu_device_to_host(filled_scaled, a->get_gpu_data(), n)) rc = 3;
            if (_gpu_device_to_host(copied, b->get_gpu_data(), n)) rc = 3;
            a->free_gpu_data();
            if (a->get_gpu_data() || a->is_on_gpu()) rc = 4;
            a->allocate_gpu_data();                      // a fresh device buffer for the same block
            if (!a->get_gpu_data() || !a->is_on_gpu()) rc = 4;
        }
        delete a;   // ~Block frees the device buffers through _gpu_free
        delete b;
    });
    return g ? 1 : rc;
}
#endif

// ---- .dat files through the reference's SetupReader ----
// Text dump (operator<< of SetupReader) of a setup file into buf; returns the length needed (excluding NUL) or -1.
long long aces4ref_setup_dump(const char* path, char* buf, long long buflen) {
    long long need = -1;
    guarded([&] {
        setup::BinaryInputFile in(path);
        if (!in.is_open()) throw std::runtime_error(std::string("cannot open ") + path);
        setup::SetupReader reader(in);
        std::ostringstream os;
        os << reader;
        const std::string s = os.str();
        need = (long long)s.size();
        if (buf && buflen > 0) {
            const size_t n = std::min((size_t)(buflen - 1), s.size());
            std::memcpy(buf, s.data(), n);
            buf[n] = 0;
        }
    });
    return need;
}

// segment extents of one index type (1001 ao, 1002 mo, 1003 moa, 1004 mob) as SetupReader::read_segment_sizes stored them;
// returns their number (-1: error, 0: the file has no table for this type)
int aces4ref_setup_segments(const char* path, int index_type, int* extents, int cap) {
    int n = -1;
    guarded([&] {
        setup::BinaryInputFile in(path);
        if (!in.is_open()) throw std::runtime_error(std::string("cannot open ") + path);
        setup::SetupReader reader(in);
        setup::SetupReader::SetupSegmentInfoMap::const_iterator it = reader.segment_map_.find(sip::intToIndexType_t(index_type));
        if (it == reader.segment_map_.end()) { n = 0; return; }
        n = (int)it->second.size();
        for (int k = 0; k < n && k < cap; ++k) extents[k] = it->second[k];
    });
    return n;
}

int aces4ref_setup_predefined_int(const char* path, const char* name, int* value) {
    return guarded([&] {
        setup::BinaryInputFile in(path);
        if (!in.is_open()) throw std::runtime_error(std::string("cannot open ") + path);
        setup::SetupReader reader(in);
        *value = reader.predefined_int(name);
    });
}

int aces4ref_setup_predefined_scalar(const char* path, const char* name, double* value) {
    return guarded([&] {
        setup::BinaryInputFile in(path);
        if (!in.is_open()) throw std::runtime_error(std::string("cannot open ") + path);
        setup::SetupReader reader(in);
        *value = reader.predefined_scalar(name);
    });
}

// ---- the reference's binary stream (io_utils.cpp): what the worker checkpoint and the .dat files are written with ----
// A tiny tagged record stream: kinds[i] 0 int, 1 double, 2 string, 3 int array, 4 double array, 5 size_t.
// write: ints/doubles/strings are consumed in order from the three pools.
int aces4ref_stream_write(const char* path, int nrec, const int* kinds, const long long* ivals, const double* dvals,
                          const char* const* svals, const int* array_len) {
    return guarded([&] {
        setup::BinaryOutputFile out(path);
        size_t ii = 0, di = 0, si = 0;
        for (int r = 0; r < nrec; ++r) {
            switch (kinds[r]) {
            case 0: out.write_int((int)ivals[ii++]); break;
            case 1: out.write_double(dvals[di++]); break;
            case 2: out.write_string(svals[si++]); break;
            case 3: {
                std::vector<int> a(array_len[r] > 0 ? array_len[r] : 1);
                for (int k = 0; k < array_len[r]; ++k) a[k] = (int)ivals[ii++];
                out.write_int_array(array_len[r], &a[0]);
                break;
            }
            case 4: {
                std::vector<double> a(array_len[r] > 0 ? array_len[r] : 1);
                for (int k = 0; k < array_len[r]; ++k) a[k] = dvals[di++];
                out.write_double_array(array_len[r], &a[0]);
                break;
            }
            case 5: out.write_size_t_val((size_t)ivals[ii++]); break;
            default: throw std::invalid_argument("aces4ref_stream_write: bad kind");
            }
        }
    });
}

// read the same record kinds back; strings are returned concatenated with '\n' separators in sbuf; array lengths in array_len
int aces4ref_stream_read(const char* path, int nrec, const int* kinds, long long* ivals, double* dvals, char* sbuf,
                         long long sbuflen, int* array_len) {
    return guarded([&] {
        setup::BinaryInputFile in(path);
        if (!in.is_open()) throw std::runtime_error(std::string("cannot open ") + path);
        size_t ii = 0, di = 0;
        std::string strings;
        for (int r = 0; r < nrec; ++r) {
            array_len[r] = 0;
            switch (kinds[r]) {
            case 0: ivals[ii++] = in.read_int(); break;
            case 1: dvals[di++] = in.read_double(); break;
            case 2: strings += in.read_string(); strings += '\n'; break;
            case 3: {
                int n = 0;
                int* a = in.read_int_array(&n);
                for (int k = 0; k < n; ++k) ivals[ii++] = a[k];
                array_len[r] = n;
                delete[] a;
                break;
            }
            case 4: {
                int n = 0;
                double* a = in.read_double_array(&n);
                for (int k = 0; k < n; ++k) dvals[di++] = a[k];
                array_len[r] = n;
                delete[] a;
                break;
            }
            case 5: ivals[ii++] = (long long)in.read_size_t(); break;
            default: throw std::invalid_argument("aces4ref_stream_read: bad kind");
            }
        }
        if ((long long)strings.size() + 1 > sbuflen) throw std::length_error("aces4ref_stream_read: string buffer too small");
        std::memcpy(sbuf, strings.c_str(), strings.size() + 1);
    });
}

}  // extern "C"

// ---- link-closure stubs: symbols the compiled reference files mention on paths these entry points never take ----
int current_line() { return 0; }   // sip_interface.cpp:196 (asks the running interpreter; there is none)

namespace sip {
Interpreter* Interpreter::global_interpreter = NULL;   // interpreter.cpp:37; only subindex descriptors dereference it
// BlockId's stream output asks SipTables for names; the shim never prints array-qualified ids
std::string SipTables::array_name(int) const { return "array"; }
int SipTables::array_rank(int) const { return MAX_RANK; }
bool SipTables::is_contiguous_local(int) const { return false; }
}  // namespace sip
