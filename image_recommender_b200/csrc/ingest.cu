// Native ingest for the build half of the hot path: SQLite rows -> float32 payloads -> pinned
// staging -> (async H2D + K-pack).  Replaces the Python row loop of
//   FAISSIndexBuilderDB._batch_records  (main/create_index.py:144-158: fetchmany over the join)
//   FAISSIndexBuilderDB._process_batch  (main/create_index.py:160-189: pickle.loads per blob)
// for databases whose blobs are what the reference's extractors write
// (vector_scripts/create_vector_base.py:142-145).  Host code only; it drives the engine through
// the public C ABI (b2k_stage_*), and talks to SQLite through dlopen("libsqlite3.so.0") with the
// handful of prototypes it needs (the image ships the library but not sqlite3.h).
#include <dlfcn.h>
#include <string.h>

#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace b2k {
namespace {

// ---- strict recogniser of pickle.dumps(np.ndarray[float32, 1-D, C-contiguous], protocol=5) ----
// numpy >= 2 reduces such an array to numpy._core.numeric._frombuffer(bytearray, dtype, shape, order):
//   80 05 | 95 <frame:8> | 8c <len> "numpy._core.numeric" 94 | 8c 0b "_frombuffer" 94 93 94 |
//   28 (MARK) | 96 <n_bytes:8> <raw little-endian float32> 94 | ... dtype('f4') reduce: 8c 02 "f4" ...
//   8c 01 "<" ... | shape tuple: (4b d | 4d d:2 | 4a d:4) 85 94 | 8c 01 "C" 94 | 74 94 52 94 2e
// The checks mirror FAISSIndexBuilderDB._decode_blob's fast path byte for byte (main/create_index.py
// of this repo), so both decoders accept exactly the same blobs.
const unsigned char kMark[] = {'_', 'f', 'r', 'o', 'm', 'b', 'u', 'f', 'f', 'e', 'r', 0x94, 0x93, 0x94, 0x28, 0x96};
const unsigned char kTail[] = {0x8c, 0x01, 'C', 0x94, 0x74, 0x94, 0x52, 0x94, 0x2e};
const unsigned char kF4[] = {0x8c, 0x02, 'f', '4', 0x94};
const unsigned char kLE[] = {0x8c, 0x01, '<', 0x94};

const unsigned char* find_bytes(const unsigned char* hay, int64_t n, const unsigned char* needle, int64_t m) {
  for (int64_t i = 0; i + m <= n; ++i)
    if (memcmp(hay + i, needle, (size_t)m) == 0) return hay + i;
  return nullptr;
}

int parse_blob(const unsigned char* b, int64_t n, const float** payload, int64_t* dim) {
  if (!b || n < 32 || b[0] != 0x80 || b[1] != 0x05) return B2K_E_UNSUPPORTED;
  const int64_t head = n < 80 ? n : 80;
  const unsigned char* m = find_bytes(b, head, kMark, sizeof(kMark));
  if (!m || m == b) return B2K_E_UNSUPPORTED;
  const int64_t off = (m - b) + (int64_t)sizeof(kMark) + 8;
  if (off > n || off > 80) return B2K_E_UNSUPPORTED;      // the length field itself lies in the first 80 bytes
  uint64_t nb = 0;
  for (int i = 7; i >= 0; --i) nb = (nb << 8) | b[off - 8 + i];
  if (nb == 0 || (nb & 3) != 0 || nb > (uint64_t)(n - off)) return B2K_E_UNSUPPORTED;
  const unsigned char* tail = b + off + nb;
  const int64_t tl = n - off - (int64_t)nb;
  if (tl <= 0 || tl >= 160) return B2K_E_UNSUPPORTED;
  if (!find_bytes(tail, tl, kF4, sizeof(kF4)) || !find_bytes(tail, tl, kLE, sizeof(kLE))) return B2K_E_UNSUPPORTED;
  const uint64_t d = nb / 4;
  unsigned char shape[16];
  int sl = 0;
  if (d < 256) { shape[sl++] = 'K'; shape[sl++] = (unsigned char)d; }
  else if (d < 65536) { shape[sl++] = 'M'; shape[sl++] = (unsigned char)(d & 0xff); shape[sl++] = (unsigned char)(d >> 8); }
  else if (d < 0x80000000ull) { shape[sl++] = 'J'; for (int i = 0; i < 4; ++i) shape[sl++] = (unsigned char)((d >> (8 * i)) & 0xff); }
  else return B2K_E_UNSUPPORTED;
  shape[sl++] = 0x85; shape[sl++] = 0x94;
  memcpy(shape + sl, kTail, sizeof(kTail));
  sl += (int)sizeof(kTail);
  if (tl < sl || memcmp(tail + tl - sl, shape, (size_t)sl) != 0) return B2K_E_UNSUPPORTED;
  *payload = reinterpret_cast<const float*>(b + off);     // unaligned: callers memcpy
  *dim = (int64_t)d;
  return 0;
}

// ---- libsqlite3 through dlopen ---------------------------------------------------------------
struct Sqlite {
  void* lib = nullptr;
  int (*open_v2)(const char*, void**, int, const char*) = nullptr;
  int (*close)(void*) = nullptr;
  int (*prepare_v2)(void*, const char*, int, void**, const char**) = nullptr;
  int (*step)(void*) = nullptr;
  int (*finalize)(void*) = nullptr;
  int (*column_count)(void*) = nullptr;
  int (*column_type)(void*, int) = nullptr;
  long long (*column_int64)(void*, int) = nullptr;
  const void* (*column_blob)(void*, int) = nullptr;
  int (*column_bytes)(void*, int) = nullptr;
  const char* (*errmsg)(void*) = nullptr;
  int (*bind_int64)(void*, int, long long) = nullptr;
  int (*reset)(void*) = nullptr;
  bool load() {
    if (lib) return true;
    for (const char* name : {"libsqlite3.so.0", "libsqlite3.so"}) {
      lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (lib) break;
    }
    if (!lib) return false;
#define B2K_SYM(field, sym) *reinterpret_cast<void**>(&field) = dlsym(lib, sym); if (!field) return false
    B2K_SYM(open_v2, "sqlite3_open_v2"); B2K_SYM(close, "sqlite3_close"); B2K_SYM(prepare_v2, "sqlite3_prepare_v2");
    B2K_SYM(step, "sqlite3_step"); B2K_SYM(finalize, "sqlite3_finalize"); B2K_SYM(column_count, "sqlite3_column_count");
    B2K_SYM(column_type, "sqlite3_column_type"); B2K_SYM(column_int64, "sqlite3_column_int64");
    B2K_SYM(column_blob, "sqlite3_column_blob"); B2K_SYM(column_bytes, "sqlite3_column_bytes"); B2K_SYM(errmsg, "sqlite3_errmsg");
    B2K_SYM(bind_int64, "sqlite3_bind_int64"); B2K_SYM(reset, "sqlite3_reset");
#undef B2K_SYM
    return true;
  }
};
constexpr int kSqliteOpenNoMutex = 0x8000;
constexpr int kSqliteOpenReadonly = 1, kSqliteRow = 100, kSqliteDone = 101, kSqliteInteger = 1, kSqliteBlob = 4;

}  // namespace
}  // namespace b2k

using namespace b2k;

extern "C" {

int b2k_parse_f32_blob(const void* blob, int64_t n_bytes, const float** payload, int64_t* dim) {
  if (!payload || !dim) { set_error("parse_f32_blob: null output"); return B2K_E_INVALID; }
  *payload = nullptr; *dim = 0;
  return parse_blob(static_cast<const unsigned char*>(blob), n_bytes, payload, dim);
}

int b2k_ingest_sqlite(b2k_index* ix, const char* db_path, const char* sql, int64_t* ids_out, int64_t ids_cap,
                      int64_t* n_added) {
  if (!ix || !db_path || !sql || !ids_out || !n_added) { set_error("ingest_sqlite: bad argument"); return B2K_E_INVALID; }
  *n_added = 0;
  static Sqlite sq;
  if (!sq.load()) { set_error("ingest_sqlite: libsqlite3 is not loadable (%s)", dlerror()); return B2K_E_UNSUPPORTED; }
  int32_t dims[B2K_MAX_TABLES];
  const int n_tables = b2k_table_dims(ix, dims);
  if (b2k_stage_rows(ix) == 0) { const int r = b2k_stage_open(ix, 4096); if (r) return r; }
  const int64_t slot_rows = b2k_stage_rows(ix);

  void* db = nullptr;
  void* stmt = nullptr;
  int rc = 0;
  if (sq.open_v2(db_path, &db, kSqliteOpenReadonly, nullptr) != 0) {
    set_error("ingest_sqlite: cannot open %s: %s", db_path, db ? sq.errmsg(db) : "out of memory");
    if (db) sq.close(db);
    return B2K_E_IO;
  }
  if (sq.prepare_v2(db, sql, -1, &stmt, nullptr) != 0) {
    set_error("ingest_sqlite: %s", sq.errmsg(db));
    sq.close(db);
    return B2K_E_IO;
  }
  if (sq.column_count(stmt) != 1 + n_tables) {
    set_error("ingest_sqlite: the query returns %d columns, expected id + %d blobs", sq.column_count(stmt), n_tables);
    sq.finalize(stmt); sq.close(db);
    return B2K_E_INVALID;
  }

  float* dst[B2K_MAX_TABLES];
  int slot = 0;
  int64_t in_slot = 0, total = 0;
  auto bind_slot = [&](int s) -> int {
    int r = b2k_stage_wait(ix, s);
    for (int t = 0; t < n_tables && !r; ++t) r = b2k_stage_ptr(ix, s, t, &dst[t]);
    return r;
  };
  rc = bind_slot(slot);
  while (!rc) {
    const int st = sq.step(stmt);
    if (st == kSqliteDone) break;
    if (st != kSqliteRow) { set_error("ingest_sqlite: %s", sq.errmsg(db)); rc = B2K_E_IO; break; }
    if (total >= ids_cap) { set_error("ingest_sqlite: more than %lld rows", (long long)ids_cap); rc = B2K_E_CAPACITY; break; }
    if (sq.column_type(stmt, 0) != kSqliteInteger) { set_error("ingest_sqlite: first column is not an integer id"); rc = B2K_E_UNSUPPORTED; break; }
    const long long id = sq.column_int64(stmt, 0);
    for (int t = 0; t < n_tables; ++t) {
      const float* payload = nullptr;
      int64_t d = 0;
      const void* blob = sq.column_type(stmt, 1 + t) == kSqliteBlob ? sq.column_blob(stmt, 1 + t) : nullptr;
      const int nb = blob ? sq.column_bytes(stmt, 1 + t) : 0;
      if (parse_blob(static_cast<const unsigned char*>(blob), nb, &payload, &d) != 0 || d != dims[t]) {
        set_error("ingest_sqlite: image id %lld, table %d: blob is not a pickled 1-D float32 ndarray of %d values",
                  id, t, dims[t]);
        rc = B2K_E_UNSUPPORTED;
        break;
      }
      memcpy(dst[t] + in_slot * dims[t], payload, (size_t)d * sizeof(float));
    }
    if (rc) break;
    ids_out[total++] = id;
    if (++in_slot == slot_rows) {
      rc = b2k_stage_commit(ix, slot, in_slot);
      in_slot = 0;
      slot ^= 1;
      if (!rc) rc = bind_slot(slot);
    }
  }
  if (!rc && in_slot > 0) rc = b2k_stage_commit(ix, slot, in_slot);
  sq.finalize(stmt);
  sq.close(db);
  // rows committed before a failure stay appended (ids_out[0, *n_added) names them): the caller decides
  const int rc2 = b2k_stage_wait(ix, 0) | b2k_stage_wait(ix, 1);
  *n_added = rc ? total - in_slot : total;
  return rc ? rc : rc2;
}

int b2k_ingest_sqlite_mt(b2k_index* ix, const char* db_path, const char* sql_range, const int64_t* id_bounds,
                         int64_t n_chunks, int32_t n_threads, int64_t* ids_out, int64_t ids_cap, int64_t* n_added) {
  if (!ix || !db_path || !sql_range || !id_bounds || !ids_out || !n_added || n_chunks < 0 || n_threads < 1) {
    set_error("ingest_sqlite_mt: bad argument");
    return B2K_E_INVALID;
  }
  *n_added = 0;
  static Sqlite sq;
  static std::mutex load_mu;
  {
    std::lock_guard<std::mutex> lk(load_mu);
    if (!sq.load()) { set_error("ingest_sqlite: libsqlite3 is not loadable (%s)", dlerror()); return B2K_E_UNSUPPORTED; }
  }
  int32_t dims[B2K_MAX_TABLES];
  const int n_tables = b2k_table_dims(ix, dims);
  const int64_t slot_rows = b2k_stage_rows(ix);
  if (slot_rows == 0) { set_error("ingest_sqlite_mt: b2k_stage_open_n(idx, rows, 2 * n_threads) first"); return B2K_E_INVALID; }
  n_threads = (int32_t)std::max<int64_t>(1, std::min<int64_t>(n_threads, std::max<int64_t>(n_chunks, 1)));
  {
    float* probe = nullptr;                      // two slots per thread must exist
    if (b2k_stage_ptr(ix, 2 * n_threads - 1, 0, &probe) != 0) { set_error("ingest_sqlite_mt: fewer than %d staging slots are open", 2 * n_threads); return B2K_E_INVALID; }
  }

  struct Shared {
    std::mutex mu;
    std::condition_variable cv;
    int64_t turn = 0;            // next chunk to commit
    int64_t total = 0;           // rows committed (== appended, in order)
    int rc = 0;                  // first failure
    std::string err;
  } sh;

  auto fail = [&](int rc, const std::string& msg) {
    std::lock_guard<std::mutex> lk(sh.mu);
    if (!sh.rc) { sh.rc = rc; sh.err = msg; }
    sh.cv.notify_all();
  };

  auto worker = [&](int p) {
    void* db = nullptr;
    void* stmt = nullptr;
    if (sq.open_v2(db_path, &db, kSqliteOpenReadonly | kSqliteOpenNoMutex, nullptr) != 0) {
      fail(B2K_E_IO, std::string("ingest_sqlite: cannot open ") + db_path + ": " + (db ? sq.errmsg(db) : "out of memory"));
      if (db) sq.close(db);
      return;
    }
    if (sq.prepare_v2(db, sql_range, -1, &stmt, nullptr) != 0) {
      fail(B2K_E_IO, std::string("ingest_sqlite: ") + sq.errmsg(db));
      sq.close(db);
      return;
    }
    if (sq.column_count(stmt) != 1 + n_tables) {
      fail(B2K_E_INVALID, "ingest_sqlite: the query returns " + std::to_string(sq.column_count(stmt)) + " columns, expected id + " +
                              std::to_string(n_tables) + " blobs");
      sq.finalize(stmt); sq.close(db);
      return;
    }
    std::vector<int64_t> ids((size_t)slot_rows);
    float* dst[B2K_MAX_TABLES];
    int64_t it = 0;
    for (int64_t c = p; c < n_chunks; c += n_threads, ++it) {
      { std::lock_guard<std::mutex> lk(sh.mu); if (sh.rc) break; }
      const int slot = 2 * p + (int)(it & 1);
      int rc = b2k_stage_wait(ix, slot);
      for (int t = 0; t < n_tables && !rc; ++t) rc = b2k_stage_ptr(ix, slot, t, &dst[t]);
      if (rc) { fail(rc, b2k_last_error()); break; }
      sq.reset(stmt);
      sq.bind_int64(stmt, 1, (long long)id_bounds[c]);
      sq.bind_int64(stmt, 2, (long long)id_bounds[c + 1]);
      int64_t n = 0;
      std::string msg;
      for (;;) {
        const int st = sq.step(stmt);
        if (st == kSqliteDone) break;
        if (st != kSqliteRow) { rc = B2K_E_IO; msg = std::string("ingest_sqlite: ") + sq.errmsg(db); break; }
        if (n >= slot_rows) { rc = B2K_E_INVALID; msg = "ingest_sqlite_mt: a chunk holds more than " + std::to_string(slot_rows) + " rows"; break; }
        if (sq.column_type(stmt, 0) != kSqliteInteger) { rc = B2K_E_UNSUPPORTED; msg = "ingest_sqlite: first column is not an integer id"; break; }
        const long long id = sq.column_int64(stmt, 0);
        for (int t = 0; t < n_tables; ++t) {
          const float* payload = nullptr;
          int64_t d = 0;
          const void* blob = sq.column_type(stmt, 1 + t) == kSqliteBlob ? sq.column_blob(stmt, 1 + t) : nullptr;
          const int nb = blob ? sq.column_bytes(stmt, 1 + t) : 0;
          if (parse_blob(static_cast<const unsigned char*>(blob), nb, &payload, &d) != 0 || d != dims[t]) {
            rc = B2K_E_UNSUPPORTED;
            msg = "ingest_sqlite: image id " + std::to_string(id) + ", table " + std::to_string(t) +
                  ": blob is not a pickled 1-D float32 ndarray of " + std::to_string(dims[t]) + " values";
            break;
          }
          memcpy(dst[t] + n * dims[t], payload, (size_t)d * sizeof(float));
        }
        if (rc) break;
        ids[(size_t)n++] = id;
      }
      if (rc) { fail(rc, msg); break; }
      // commit in chunk order: rows, offsets and ids_out are those of the single-threaded loop
      std::unique_lock<std::mutex> lk(sh.mu);
      sh.cv.wait(lk, [&] { return sh.turn == c || sh.rc != 0; });
      if (sh.rc) break;
      if (sh.total + n > ids_cap) {
        sh.rc = B2K_E_CAPACITY; sh.err = "ingest_sqlite: more than " + std::to_string(ids_cap) + " rows";
        sh.cv.notify_all();
        break;
      }
      rc = b2k_stage_commit(ix, slot, n);
      if (rc) { sh.rc = rc; sh.err = b2k_last_error(); sh.cv.notify_all(); break; }
      memcpy(ids_out + sh.total, ids.data(), (size_t)n * sizeof(int64_t));
      sh.total += n;
      sh.turn = c + 1;
      sh.cv.notify_all();
    }
    sq.finalize(stmt);
    sq.close(db);
  };

  std::vector<std::thread> th;
  for (int p = 0; p < n_threads; ++p) th.emplace_back(worker, p);
  for (std::thread& t : th) t.join();
  int rc2 = 0;
  for (int s = 0; s < 2 * n_threads; ++s) rc2 |= b2k_stage_wait(ix, s);
  *n_added = sh.total;
  if (sh.rc) { set_error("%s", sh.err.c_str()); return sh.rc; }
  return rc2;
}

}  // extern "C"
