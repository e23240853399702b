// edl_kv_server -- native (C++17, epoll, single thread) build of the coordination store.
//
// The reference keeps every piece of cluster state in an external etcd v3 daemon
// (python/edl/utils/constants.py:15-39, python/edl/discovery/etcd_client.py:51-263).  This is the same store as
// edl_b200/store/kv_server.py -- identical wire protocol (4-byte big-endian length + msgpack map), identical
// semantics (global revision, per-key create/mod revision + version, leases with TTL / keep-alive / revoke,
// put-if-absent, compare-and-swap transactions, sorted prefix ranges with the header revision, prefix watches
// replayed from a start revision) and the same snapshot file, so either server can take over the other's data
// directory and the Python client (store/client.py) cannot tell them apart.  One epoll loop owns all state:
// no locks, requests of a connection are answered in order, watch events are queued on the watcher's
// connection before the response of the request that caused them.
//
//   edl_kv_server --host 0.0.0.0 --port 2379 [--data_dir DIR] [--snapshot_interval 2.0]
//
// Prints "listening on HOST:PORT" on stdout once the socket is bound (port 0 = pick a free one).
#include <arpa/inet.h>
#include <errno.h>
#include <fcntl.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <signal.h>
#include <sys/epoll.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace {

// ------------------------------------------------------------------------------------------ msgpack
struct Value {
  enum Type { NIL, BOOL, INT, FLOAT, STR, BIN, ARRAY, MAP } type = NIL;
  bool b = false;
  int64_t i = 0;
  double d = 0.0;
  std::string s;                                     // STR / BIN payload
  std::vector<Value> arr;                            // ARRAY
  std::vector<std::pair<std::string, Value>> map;    // MAP (string keys only: that is all the protocol uses)

  static Value nil() { return Value(); }
  static Value boolean(bool v) { Value x; x.type = BOOL; x.b = v; return x; }
  static Value integer(int64_t v) { Value x; x.type = INT; x.i = v; return x; }
  static Value real(double v) { Value x; x.type = FLOAT; x.d = v; return x; }
  static Value str(std::string v) { Value x; x.type = STR; x.s = std::move(v); return x; }
  static Value bin(std::string v) { Value x; x.type = BIN; x.s = std::move(v); return x; }
  static Value array() { Value x; x.type = ARRAY; return x; }
  static Value object() { Value x; x.type = MAP; return x; }

  Value& set(const std::string& k, Value v) {
    map.emplace_back(k, std::move(v));
    return *this;
  }
  const Value* get(const char* k) const {
    if (type != MAP) return nullptr;
    for (const auto& kv : map)
      if (kv.first == k) return &kv.second;
    return nullptr;
  }
  bool truthy() const {
    switch (type) {
      case NIL: return false;
      case BOOL: return b;
      case INT: return i != 0;
      case FLOAT: return d != 0.0;
      case STR: case BIN: return !s.empty();
      case ARRAY: return !arr.empty();
      case MAP: return !map.empty();
    }
    return false;
  }
  int64_t as_int(int64_t dflt = 0) const {
    if (type == INT) return i;
    if (type == FLOAT) return (int64_t)d;
    if (type == BOOL) return b ? 1 : 0;
    return dflt;
  }
  double as_float(double dflt = 0.0) const {
    if (type == FLOAT) return d;
    if (type == INT) return (double)i;
    return dflt;
  }
};

struct DecodeError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

class Decoder {
 public:
  Decoder(const uint8_t* p, size_t n) : p_(p), end_(p + n) {}
  Value decode(int depth = 0) {
    if (depth > 64) throw DecodeError("nesting too deep");
    const uint8_t t = u8();
    if (t <= 0x7f) return Value::integer(t);
    if (t >= 0xe0) return Value::integer((int8_t)t);
    if ((t & 0xe0) == 0xa0) return Value::str(bytes(t & 0x1f));
    if ((t & 0xf0) == 0x90) return array(t & 0x0f, depth);
    if ((t & 0xf0) == 0x80) return object(t & 0x0f, depth);
    switch (t) {
      case 0xc0: return Value::nil();
      case 0xc2: return Value::boolean(false);
      case 0xc3: return Value::boolean(true);
      case 0xc4: return Value::bin(bytes(u8()));
      case 0xc5: return Value::bin(bytes(u16()));
      case 0xc6: return Value::bin(bytes(u32()));
      case 0xca: { uint32_t v = u32(); float f; memcpy(&f, &v, 4); return Value::real(f); }
      case 0xcb: { uint64_t v = u64(); double f; memcpy(&f, &v, 8); return Value::real(f); }
      case 0xcc: return Value::integer(u8());
      case 0xcd: return Value::integer(u16());
      case 0xce: return Value::integer(u32());
      case 0xcf: return Value::integer((int64_t)u64());
      case 0xd0: return Value::integer((int8_t)u8());
      case 0xd1: return Value::integer((int16_t)u16());
      case 0xd2: return Value::integer((int32_t)u32());
      case 0xd3: return Value::integer((int64_t)u64());
      case 0xd9: return Value::str(bytes(u8()));
      case 0xda: return Value::str(bytes(u16()));
      case 0xdb: return Value::str(bytes(u32()));
      case 0xdc: return array(u16(), depth);
      case 0xdd: return array(u32(), depth);
      case 0xde: return object(u16(), depth);
      case 0xdf: return object(u32(), depth);
      default: throw DecodeError("unsupported msgpack type");
    }
  }

 private:
  const uint8_t* p_;
  const uint8_t* end_;
  void need(size_t n) {
    if ((size_t)(end_ - p_) < n) throw DecodeError("truncated message");
  }
  uint8_t u8() { need(1); return *p_++; }
  uint16_t u16() { need(2); uint16_t v = (uint16_t)(p_[0] << 8 | p_[1]); p_ += 2; return v; }
  uint32_t u32() {
    need(4);
    uint32_t v = (uint32_t)p_[0] << 24 | (uint32_t)p_[1] << 16 | (uint32_t)p_[2] << 8 | p_[3];
    p_ += 4;
    return v;
  }
  uint64_t u64() { uint64_t hi = u32(); return hi << 32 | u32(); }
  std::string bytes(size_t n) {
    need(n);
    std::string s(reinterpret_cast<const char*>(p_), n);
    p_ += n;
    return s;
  }
  Value array(size_t n, int depth) {
    Value v = Value::array();
    v.arr.reserve(std::min<size_t>(n, 4096));
    for (size_t k = 0; k < n; ++k) v.arr.push_back(decode(depth + 1));
    return v;
  }
  Value object(size_t n, int depth) {
    Value v = Value::object();
    for (size_t k = 0; k < n; ++k) {
      Value key = decode(depth + 1);
      if (key.type != Value::STR && key.type != Value::BIN) throw DecodeError("map key is not a string");
      Value val = decode(depth + 1);
      v.map.emplace_back(std::move(key.s), std::move(val));
    }
    return v;
  }
};

void put_be(std::string& out, uint64_t v, int nbytes) {
  for (int k = nbytes - 1; k >= 0; --k) out.push_back((char)(v >> (8 * k)));
}

void encode(const Value& v, std::string& out) {
  switch (v.type) {
    case Value::NIL: out.push_back((char)0xc0); break;
    case Value::BOOL: out.push_back((char)(v.b ? 0xc3 : 0xc2)); break;
    case Value::INT:
      if (v.i >= 0) {
        if (v.i <= 0x7f) out.push_back((char)v.i);
        else if (v.i <= 0xff) { out.push_back((char)0xcc); put_be(out, v.i, 1); }
        else if (v.i <= 0xffff) { out.push_back((char)0xcd); put_be(out, v.i, 2); }
        else if (v.i <= 0xffffffffLL) { out.push_back((char)0xce); put_be(out, v.i, 4); }
        else { out.push_back((char)0xcf); put_be(out, v.i, 8); }
      } else {
        if (v.i >= -32) out.push_back((char)v.i);
        else if (v.i >= -128) { out.push_back((char)0xd0); put_be(out, (uint8_t)v.i, 1); }
        else if (v.i >= -32768) { out.push_back((char)0xd1); put_be(out, (uint16_t)v.i, 2); }
        else if (v.i >= -2147483648LL) { out.push_back((char)0xd2); put_be(out, (uint32_t)v.i, 4); }
        else { out.push_back((char)0xd3); put_be(out, (uint64_t)v.i, 8); }
      }
      break;
    case Value::FLOAT: {
      uint64_t bits;
      memcpy(&bits, &v.d, 8);
      out.push_back((char)0xcb);
      put_be(out, bits, 8);
      break;
    }
    case Value::STR:
      if (v.s.size() <= 31) out.push_back((char)(0xa0 | v.s.size()));
      else if (v.s.size() <= 0xff) { out.push_back((char)0xd9); put_be(out, v.s.size(), 1); }
      else if (v.s.size() <= 0xffff) { out.push_back((char)0xda); put_be(out, v.s.size(), 2); }
      else { out.push_back((char)0xdb); put_be(out, v.s.size(), 4); }
      out += v.s;
      break;
    case Value::BIN:
      if (v.s.size() <= 0xff) { out.push_back((char)0xc4); put_be(out, v.s.size(), 1); }
      else if (v.s.size() <= 0xffff) { out.push_back((char)0xc5); put_be(out, v.s.size(), 2); }
      else { out.push_back((char)0xc6); put_be(out, v.s.size(), 4); }
      out += v.s;
      break;
    case Value::ARRAY:
      if (v.arr.size() <= 15) out.push_back((char)(0x90 | v.arr.size()));
      else if (v.arr.size() <= 0xffff) { out.push_back((char)0xdc); put_be(out, v.arr.size(), 2); }
      else { out.push_back((char)0xdd); put_be(out, v.arr.size(), 4); }
      for (const auto& e : v.arr) encode(e, out);
      break;
    case Value::MAP:
      if (v.map.size() <= 15) out.push_back((char)(0x80 | v.map.size()));
      else if (v.map.size() <= 0xffff) { out.push_back((char)0xde); put_be(out, v.map.size(), 2); }
      else { out.push_back((char)0xdf); put_be(out, v.map.size(), 4); }
      for (const auto& kv : v.map) {
        encode(Value::str(kv.first), out);
        encode(kv.second, out);
      }
      break;
  }
}

// ------------------------------------------------------------------------------------------ state
using Clock = std::chrono::steady_clock;

double now_s() { return std::chrono::duration<double>(Clock::now().time_since_epoch()).count(); }

struct KeyValue {
  std::string value;
  int64_t create_rev = 0, mod_rev = 0, version = 0, lease = 0;
};

struct Lease {
  int64_t id = 0;
  double ttl = 0, expiry = 0;
  std::set<std::string> keys;
};

struct Event {
  int64_t rev;
  Value wire;        // {"type", "key", "kv", "revision"}
  std::string key;
};

struct Conn {
  int fd = -1;
  uint64_t cid = 0;
  std::string in, out;
  size_t out_off = 0;          // bytes of `out` already sent (a slow watcher must not cost a memmove per send)
  bool want_write = false, dead = false;
};

struct Watcher {
  Conn* conn;
  int64_t wid;
  std::string prefix;
  bool has_end = false;
  std::string end;
  bool matches(const std::string& key) const {
    return key.compare(0, prefix.size(), prefix) == 0 && key.size() >= prefix.size() && (!has_end || key < end);
  }
};

struct RequestError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

Value kv_wire(const std::string& key, const KeyValue& kv) {
  Value w = Value::object();
  w.set("key", Value::str(key)).set("value", Value::bin(kv.value)).set("create_revision", Value::integer(kv.create_rev))
      .set("mod_revision", Value::integer(kv.mod_rev)).set("version", Value::integer(kv.version))
      .set("lease", Value::integer(kv.lease));
  return w;
}

constexpr size_t kMaxBacklogBytes = 64u << 20;   // a client that does not read its messages is dropped, not buffered forever

void queue_msg(Conn* c, const Value& msg) {
  if (c->dead) return;
  if (c->out.size() - c->out_off > kMaxBacklogBytes) {
    c->dead = true;            // it reconnects and re-watches from its last seen revision
    c->out.clear();
    c->out_off = 0;
    return;
  }
  std::string body;
  encode(msg, body);
  put_be(c->out, body.size(), 4);
  c->out += body;
}

class Store {
 public:
  explicit Store(size_t history = 100000) : history_(history) {
    next_lease_ = (int64_t)(std::chrono::duration_cast<std::chrono::milliseconds>(
                                std::chrono::system_clock::now().time_since_epoch()).count() % (1LL << 30)) + 1;
  }
  int64_t rev() const { return rev_; }
  size_t keys() const { return kv_.size(); }

  Value handle(const Value& req, Conn* conn) {
    const Value* mv = req.get("method");
    const std::string m = mv != nullptr ? mv->s : "";
    Value r = Value::object();
    if (m == "put") {
      const std::string& key = need_str(req, "key");
      const Value* ine = req.get("if_not_exists");
      if (ine != nullptr && ine->truthy() && kv_.count(key)) {
        r.set("ok", Value::boolean(true)).set("succeeded", Value::boolean(false)).set("revision", Value::integer(rev_));
        return r;
      }
      Value wire = put(key, opt_bytes(req, "value"), opt_int(req, "lease"));
      r.set("ok", Value::boolean(true)).set("succeeded", Value::boolean(true)).set("revision", Value::integer(rev_))
          .set("kv", std::move(wire));
      return r;
    }
    if (m == "get" || m == "delete") {
      r.set("ok", Value::boolean(true)).set("revision", Value::integer(rev_));    // header revision before the op
      Value res = apply(m, req);
      for (auto& kv : res.map) r.set(kv.first, std::move(kv.second));
      return r;
    }
    if (m == "txn") {
      bool ok = true;
      if (const Value* cmp = req.get("compare"))
        for (const auto& c : cmp->arr) ok = compare(c) && ok;
      Value results = Value::array();
      const Value* ops = req.get(ok ? "success" : "failure");
      if (ops != nullptr)
        for (const auto& op : ops->arr) {
          const Value* t = op.get("op");
          results.arr.push_back(apply(t != nullptr ? t->s : "", op));
        }
      r.set("ok", Value::boolean(true)).set("succeeded", Value::boolean(ok)).set("revision", Value::integer(rev_))
          .set("results", std::move(results));
      return r;
    }
    if (m == "lease_grant") {
      int64_t lid = opt_int(req, "lease_id");
      if (lid == 0) lid = next_lease_;
      next_lease_ = std::max(next_lease_, lid) + 1;
      const Value* ttl = req.get("ttl");
      if (ttl == nullptr) throw RequestError("KeyError: 'ttl'");
      Lease le;
      le.id = lid;
      le.ttl = ttl->as_float();
      le.expiry = now_s() + le.ttl;
      leases_[lid] = std::move(le);
      r.set("ok", Value::boolean(true)).set("lease", Value::integer(lid)).set("ttl", Value::real(ttl->as_float()));
      return r;
    }
    if (m == "lease_keepalive") {
      auto it = leases_.find(opt_int(req, "lease"));
      r.set("ok", Value::boolean(true));
      if (it == leases_.end()) {
        r.set("ttl", Value::integer(0));
      } else {
        it->second.expiry = now_s() + it->second.ttl;
        r.set("ttl", Value::real(it->second.ttl));
      }
      return r;
    }
    if (m == "lease_revoke") {
      revoke(opt_int(req, "lease"));
      r.set("ok", Value::boolean(true));
      return r;
    }
    if (m == "lease_ttl") {
      auto it = leases_.find(opt_int(req, "lease"));
      Value keys = Value::array();
      r.set("ok", Value::boolean(true));
      if (it == leases_.end()) {
        r.set("ttl", Value::integer(-1));
      } else {
        r.set("ttl", Value::real(std::max(0.0, it->second.expiry - now_s())));
        for (const auto& k : it->second.keys) keys.arr.push_back(Value::str(k));
      }
      r.set("keys", std::move(keys));
      return r;
    }
    if (m == "watch") {
      Watcher w;
      w.conn = conn;
      w.wid = opt_int(req, "watch_id");
      w.prefix = need_str(req, "key");
      if (const Value* e = req.get("end"))
        if (e->type == Value::STR || e->type == Value::BIN) { w.has_end = true; w.end = e->s; }
      const int64_t start = opt_int(req, "start_revision");
      if (start != 0) {
        Value backlog = Value::array();
        for (const auto& ev : events_)
          if (ev.rev >= start && w.matches(ev.key)) backlog.arr.push_back(ev.wire);
        if (!backlog.arr.empty()) push(w, std::move(backlog));
      }
      watchers_[{conn->cid, w.wid}] = std::move(w);
      r.set("ok", Value::boolean(true)).set("watch_id", Value::integer(opt_int(req, "watch_id")))
          .set("revision", Value::integer(rev_));
      return r;
    }
    if (m == "cancel_watch") {
      watchers_.erase({conn->cid, opt_int(req, "watch_id")});
      r.set("ok", Value::boolean(true));
      return r;
    }
    if (m == "status") {
      r.set("ok", Value::boolean(true)).set("revision", Value::integer(rev_)).set("keys", Value::integer((int64_t)kv_.size()))
          .set("leases", Value::integer((int64_t)leases_.size())).set("server", Value::str("native"));
      return r;
    }
    r.set("ok", Value::boolean(false)).set("error", Value::str("unknown method '" + m + "'"));
    return r;
  }

  void expire() {
    const double now = now_s();
    std::vector<int64_t> dead;
    for (const auto& kv : leases_)
      if (kv.second.expiry <= now) dead.push_back(kv.first);
    for (int64_t lid : dead) revoke(lid);
  }

  void drop_conn(uint64_t cid) {
    for (auto it = watchers_.begin(); it != watchers_.end();)
      it = it->first.first == cid ? watchers_.erase(it) : std::next(it);
  }

  // -- durability: the same file edl_b200/store/kv_server.py writes ------------------------------------
  std::string snapshot() const {
    Value d = Value::object();
    Value kvs = Value::array(), leases = Value::array();
    for (const auto& kv : kv_) {
      Value e = Value::array();
      e.arr = {Value::str(kv.first), Value::bin(kv.second.value), Value::integer(kv.second.create_rev),
               Value::integer(kv.second.mod_rev), Value::integer(kv.second.version), Value::integer(kv.second.lease)};
      kvs.arr.push_back(std::move(e));
    }
    const double now = now_s();
    for (const auto& le : leases_) {
      Value e = Value::array();
      e.arr = {Value::integer(le.second.id), Value::real(le.second.ttl), Value::real(std::max(0.0, le.second.expiry - now))};
      leases.arr.push_back(std::move(e));
    }
    d.set("rev", Value::integer(rev_)).set("next_lease", Value::integer(next_lease_)).set("kv", std::move(kvs))
        .set("leases", std::move(leases));
    std::string out;
    encode(d, out);
    return out;
  }

  void restore(const std::string& blob, double lease_grace) {
    Decoder dec(reinterpret_cast<const uint8_t*>(blob.data()), blob.size());
    const Value d = dec.decode();
    const Value *rv = d.get("rev"), *nl = d.get("next_lease"), *kvs = d.get("kv"), *ls = d.get("leases");
    if (rv == nullptr || kvs == nullptr) throw DecodeError("not a snapshot");
    rev_ = rv->as_int(1);
    if (nl != nullptr) next_lease_ = std::max(next_lease_, nl->as_int());
    kv_.clear();
    leases_.clear();
    const double now = now_s();
    if (ls != nullptr)
      for (const auto& e : ls->arr) {
        if (e.arr.size() < 3) continue;
        Lease le;
        le.id = e.arr[0].as_int();
        le.ttl = e.arr[1].as_float();
        le.expiry = now + std::max(e.arr[2].as_float(), lease_grace);
        leases_[le.id] = std::move(le);
      }
    for (const auto& e : kvs->arr) {
      if (e.arr.size() < 6) continue;
      KeyValue kv;
      kv.value = e.arr[1].s;
      kv.create_rev = e.arr[2].as_int();
      kv.mod_rev = e.arr[3].as_int();
      kv.version = e.arr[4].as_int();
      kv.lease = e.arr[5].as_int();
      if (kv.lease != 0) {
        auto it = leases_.find(kv.lease);
        if (it != leases_.end()) it->second.keys.insert(e.arr[0].s);
        else kv.lease = 0;
      }
      kv_[e.arr[0].s] = std::move(kv);
    }
  }

 private:
  std::map<std::string, KeyValue> kv_;                 // ordered: prefix ranges come out sorted
  std::unordered_map<int64_t, Lease> leases_;
  std::deque<Event> events_;
  std::map<std::pair<uint64_t, int64_t>, Watcher> watchers_;
  int64_t rev_ = 1, next_lease_ = 1;
  size_t history_;

  static const std::string& need_str(const Value& req, const char* k) {
    const Value* v = req.get(k);
    if (v == nullptr || (v->type != Value::STR && v->type != Value::BIN))
      throw RequestError(std::string("KeyError: '") + k + "'");
    return v->s;
  }
  static std::string opt_bytes(const Value& req, const char* k) {
    const Value* v = req.get(k);
    return v != nullptr && (v->type == Value::STR || v->type == Value::BIN) ? v->s : std::string();
  }
  static int64_t opt_int(const Value& req, const char* k) {
    const Value* v = req.get(k);
    return v != nullptr ? v->as_int() : 0;
  }

  void push(const Watcher& w, Value events) {
    Value msg = Value::object();
    msg.set("watch_id", Value::integer(w.wid)).set("events", std::move(events)).set("revision", Value::integer(rev_));
    queue_msg(w.conn, msg);
    w.conn->want_write = true;
  }

  void emit(const char* type, const std::string& key, Value wire) {
    Event ev;
    ev.rev = rev_;
    ev.key = key;
    ev.wire = Value::object();
    ev.wire.set("type", Value::str(type)).set("key", Value::str(key)).set("kv", std::move(wire))
        .set("revision", Value::integer(rev_));
    for (const auto& w : watchers_)
      if (w.second.matches(key)) {
        Value one = Value::array();
        one.arr.push_back(ev.wire);
        push(w.second, std::move(one));
      }
    events_.push_back(std::move(ev));
    if (events_.size() > history_) events_.pop_front();
  }

  Value put(const std::string& key, std::string value, int64_t lease) {
    if (lease != 0 && !leases_.count(lease))
      throw RequestError("KeyError: 'lease " + std::to_string(lease) + " not found'");
    ++rev_;
    auto it = kv_.find(key);
    KeyValue cur;
    if (it != kv_.end()) {
      const KeyValue& old = it->second;
      if (old.lease != 0 && old.lease != lease) {
        auto lo = leases_.find(old.lease);
        if (lo != leases_.end()) lo->second.keys.erase(key);
      }
      cur.create_rev = old.create_rev;
      cur.version = old.version + 1;
    } else {
      cur.create_rev = rev_;
      cur.version = 1;
    }
    cur.value = std::move(value);
    cur.mod_rev = rev_;
    cur.lease = lease;
    if (lease != 0) leases_[lease].keys.insert(key);
    Value wire = kv_wire(key, cur);
    kv_[key] = std::move(cur);
    emit("put", key, wire);
    return wire;
  }

  int64_t del(const std::string& key) {
    auto it = kv_.find(key);
    if (it == kv_.end()) return 0;
    const int64_t lease = it->second.lease;
    kv_.erase(it);
    ++rev_;
    if (lease != 0) {
      auto lo = leases_.find(lease);
      if (lo != leases_.end()) lo->second.keys.erase(key);
    }
    KeyValue tomb;
    tomb.mod_rev = rev_;
    emit("delete", key, kv_wire(key, tomb));
    return 1;
  }

  Value range(const std::string& prefix, const Value* end) const {
    Value out = Value::array();
    const bool has_end = end != nullptr && (end->type == Value::STR || end->type == Value::BIN);
    for (auto it = kv_.lower_bound(prefix); it != kv_.end(); ++it) {
      if (it->first.compare(0, prefix.size(), prefix) != 0) break;
      if (has_end && !(it->first < end->s)) continue;
      out.arr.push_back(kv_wire(it->first, it->second));
    }
    return out;
  }

  bool compare(const Value& c) const {
    const std::string& key = need_str(c, "key");
    const Value* tv = c.get("target");
    const Value* ov = c.get("op");
    const std::string target = tv != nullptr ? tv->s : "value";
    const std::string op = ov != nullptr ? ov->s : "==";
    auto it = kv_.find(key);
    const KeyValue* kv = it != kv_.end() ? &it->second : nullptr;
    int cmp;
    if (target == "version" || target == "create" || target == "mod") {
      const int64_t lhs = kv == nullptr ? 0 : (target == "version" ? kv->version : target == "create" ? kv->create_rev : kv->mod_rev);
      const int64_t rhs = opt_int(c, "value");
      cmp = lhs < rhs ? -1 : (lhs > rhs ? 1 : 0);
    } else {
      if (kv == nullptr) return op == "!=";
      const int r = kv->value.compare(opt_bytes(c, "value"));
      cmp = r < 0 ? -1 : (r > 0 ? 1 : 0);
    }
    if (op == "==") return cmp == 0;
    if (op == "!=") return cmp != 0;
    if (op == ">") return cmp > 0;
    if (op == "<") return cmp < 0;
    throw RequestError("KeyError: '" + op + "'");
  }

  Value apply(const std::string& t, const Value& op) {
    Value r = Value::object();
    if (t == "put") {
      r.set("put", put(need_str(op, "key"), opt_bytes(op, "value"), opt_int(op, "lease")));
      return r;
    }
    const Value* pv = op.get("prefix");
    const bool prefix = pv != nullptr && pv->truthy();
    if (t == "delete") {
      int64_t n = 0;
      const std::string& key = need_str(op, "key");
      if (prefix) {
        std::vector<std::string> doomed;
        for (auto it = kv_.lower_bound(key); it != kv_.end() && it->first.compare(0, key.size(), key) == 0; ++it)
          doomed.push_back(it->first);
        for (const auto& k : doomed) n += del(k);
      } else {
        n = del(key);
      }
      r.set("deleted", Value::integer(n));
      return r;
    }
    if (t == "get") {
      const std::string& key = need_str(op, "key");
      if (prefix) {
        r.set("kvs", range(key, op.get("end")));
      } else {
        Value kvs = Value::array();
        auto it = kv_.find(key);
        if (it != kv_.end()) kvs.arr.push_back(kv_wire(key, it->second));
        r.set("kvs", std::move(kvs));
      }
      return r;
    }
    throw RequestError("ValueError: unknown op '" + t + "'");
  }

  void revoke(int64_t lid) {
    auto it = leases_.find(lid);
    if (it == leases_.end()) return;
    const std::vector<std::string> keys(it->second.keys.begin(), it->second.keys.end());
    leases_.erase(it);
    for (const auto& k : keys) del(k);
  }
};

// ------------------------------------------------------------------------------------------ server
volatile sig_atomic_t g_stop = 0;
void on_signal(int) { g_stop = 1; }

bool write_snapshot(const Store& store, const std::string& dir) {
  const std::string blob = store.snapshot();
  const std::string tmp = dir + "/snapshot.bin.tmp", fin = dir + "/snapshot.bin";
  FILE* f = fopen(tmp.c_str(), "wb");
  if (f == nullptr) return false;
  const bool ok = fwrite(blob.data(), 1, blob.size(), f) == blob.size() && fflush(f) == 0 && fsync(fileno(f)) == 0;
  fclose(f);
  return ok && rename(tmp.c_str(), fin.c_str()) == 0;     // atomic: readers never see a torn file
}

void set_nonblock(int fd) { fcntl(fd, F_SETFL, fcntl(fd, F_GETFL, 0) | O_NONBLOCK); }

constexpr size_t kMaxMessage = 256u << 20;

}  // namespace

int main(int argc, char** argv) {
  std::string host = "0.0.0.0", data_dir;
  int port = 2379;
  double snapshot_interval = 2.0;
  for (int a = 1; a < argc; ++a) {
    const std::string k = argv[a];
    const char* v = a + 1 < argc ? argv[a + 1] : nullptr;
    if (k == "--host" && v) { host = v; ++a; }
    else if (k == "--port" && v) { port = atoi(v); ++a; }
    else if (k == "--data_dir" && v) { data_dir = v; ++a; }
    else if (k == "--snapshot_interval" && v) { snapshot_interval = atof(v); ++a; }
    else if (k == "--log_level" && v) { ++a; }
    else if (k == "-h" || k == "--help") {
      printf("usage: edl_kv_server [--host H] [--port P] [--data_dir DIR] [--snapshot_interval S]\n");
      return 0;
    }
  }
  signal(SIGPIPE, SIG_IGN);
  struct sigaction sa {};
  sa.sa_handler = on_signal;
  sigaction(SIGTERM, &sa, nullptr);
  sigaction(SIGINT, &sa, nullptr);

  Store store;
  int64_t snap_rev = -1;
  if (!data_dir.empty()) {
    mkdir(data_dir.c_str(), 0755);
    const std::string path = data_dir + "/snapshot.bin";
    if (FILE* f = fopen(path.c_str(), "rb")) {
      std::string blob;
      char buf[65536];
      size_t n;
      while ((n = fread(buf, 1, sizeof(buf), f)) > 0) blob.append(buf, n);
      fclose(f);
      try {
        store.restore(blob, 10.0);
        snap_rev = store.rev();
        fprintf(stderr, "restored %zu keys at revision %lld from %s\n", store.keys(), (long long)store.rev(), path.c_str());
      } catch (const std::exception& e) {
        fprintf(stderr, "ignoring unreadable snapshot %s: %s\n", path.c_str(), e.what());
      }
    }
  }

  const int lfd = socket(AF_INET, SOCK_STREAM, 0);
  int one = 1;
  setsockopt(lfd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
  sockaddr_in addr {};
  addr.sin_family = AF_INET;
  addr.sin_port = htons((uint16_t)port);
  if (inet_pton(AF_INET, host.c_str(), &addr.sin_addr) != 1) addr.sin_addr.s_addr = htonl(INADDR_ANY);
  if (bind(lfd, reinterpret_cast<sockaddr*>(&addr), sizeof(addr)) != 0 || listen(lfd, 256) != 0) {
    perror("bind/listen");
    return 1;
  }
  socklen_t alen = sizeof(addr);
  getsockname(lfd, reinterpret_cast<sockaddr*>(&addr), &alen);
  set_nonblock(lfd);
  printf("listening on %s:%d\n", host.c_str(), (int)ntohs(addr.sin_port));
  fflush(stdout);

  const int ep = epoll_create1(0);
  epoll_event ev {};
  ev.events = EPOLLIN;
  ev.data.fd = lfd;
  epoll_ctl(ep, EPOLL_CTL_ADD, lfd, &ev);

  std::unordered_map<int, std::unique_ptr<Conn>> conns;
  uint64_t next_cid = 0;
  double last_expire = now_s(), last_snap = now_s();

  auto flush = [&](Conn* c) {
    while (c->out_off < c->out.size()) {
      const ssize_t n = send(c->fd, c->out.data() + c->out_off, c->out.size() - c->out_off, MSG_NOSIGNAL);
      if (n > 0) {
        c->out_off += (size_t)n;
      } else if (n < 0 && (errno == EAGAIN || errno == EWOULDBLOCK)) {
        break;
      } else if (n < 0 && errno == EINTR) {
        continue;
      } else {
        c->dead = true;
        c->out.clear();
        c->out_off = 0;
        return;
      }
    }
    if (c->out_off == c->out.size()) {
      c->out.clear();
      c->out_off = 0;
    } else if (c->out_off > (1u << 20) && c->out_off * 2 > c->out.size()) {
      c->out.erase(0, c->out_off);          // compact rarely, not per send
      c->out_off = 0;
    }
    epoll_event e {};
    e.events = EPOLLIN | (c->out.empty() ? 0u : (uint32_t)EPOLLOUT);
    e.data.fd = c->fd;
    epoll_ctl(ep, EPOLL_CTL_MOD, c->fd, &e);
    c->want_write = false;
  };
  auto flush_pending = [&]() {
    for (auto& kv : conns)
      if (!kv.second->dead && (kv.second->want_write || !kv.second->out.empty())) flush(kv.second.get());
  };
  auto close_conn = [&](int fd) {
    auto it = conns.find(fd);
    if (it == conns.end()) return;
    store.drop_conn(it->second->cid);
    epoll_ctl(ep, EPOLL_CTL_DEL, fd, nullptr);
    close(fd);
    conns.erase(it);
  };

  std::vector<epoll_event> events(256);
  while (!g_stop) {
    const int n = epoll_wait(ep, events.data(), (int)events.size(), 50);
    for (int k = 0; k < n; ++k) {
      const int fd = events[k].data.fd;
      if (fd == lfd) {
        for (;;) {
          const int cfd = accept(lfd, nullptr, nullptr);
          if (cfd < 0) break;
          set_nonblock(cfd);
          setsockopt(cfd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
          auto c = std::make_unique<Conn>();
          c->fd = cfd;
          c->cid = ++next_cid;
          epoll_event e {};
          e.events = EPOLLIN;
          e.data.fd = cfd;
          epoll_ctl(ep, EPOLL_CTL_ADD, cfd, &e);
          conns[cfd] = std::move(c);
        }
        continue;
      }
      auto it = conns.find(fd);
      if (it == conns.end()) continue;
      Conn* c = it->second.get();
      if (events[k].events & (EPOLLHUP | EPOLLERR)) c->dead = true;
      if (!c->dead && (events[k].events & EPOLLIN)) {
        char buf[65536];
        for (;;) {
          const ssize_t r = recv(fd, buf, sizeof(buf), 0);
          if (r > 0) {
            c->in.append(buf, (size_t)r);
          } else if (r == 0) {
            c->dead = true;
            break;
          } else if (errno == EAGAIN || errno == EWOULDBLOCK) {
            break;
          } else if (errno != EINTR) {
            c->dead = true;
            break;
          }
        }
        size_t off = 0;
        while (c->in.size() - off >= 4) {
          const uint8_t* p = reinterpret_cast<const uint8_t*>(c->in.data()) + off;
          const size_t len = (size_t)p[0] << 24 | (size_t)p[1] << 16 | (size_t)p[2] << 8 | p[3];
          if (len > kMaxMessage) { c->dead = true; break; }
          if (c->in.size() - off - 4 < len) break;
          Value resp;
          Value id = Value::nil();
          try {
            Decoder dec(p + 4, len);
            const Value req = dec.decode();
            if (const Value* i = req.get("id")) id = *i;
            try {
              resp = store.handle(req, c);
            } catch (const RequestError& e) {
              resp = Value::object();
              resp.set("ok", Value::boolean(false)).set("error", Value::str(e.what()));
            }
          } catch (const DecodeError& e) {
            resp = Value::object();
            resp.set("ok", Value::boolean(false)).set("error", Value::str(std::string("DecodeError: ") + e.what()));
          }
          resp.set("id", std::move(id));
          queue_msg(c, resp);
          off += 4 + len;
        }
        if (off > 0) c->in.erase(0, off);
      }
      if (!c->dead && (events[k].events & EPOLLOUT)) c->want_write = true;
    }
    const double now = now_s();
    if (now - last_expire >= 0.1) {
      store.expire();
      last_expire = now;
    }
    flush_pending();
    std::vector<int> dead;
    for (auto& kv : conns)
      if (kv.second->dead) dead.push_back(kv.first);
    for (int fd : dead) close_conn(fd);
    if (!data_dir.empty() && now - last_snap >= snapshot_interval) {
      last_snap = now;
      if (store.rev() != snap_rev && write_snapshot(store, data_dir)) snap_rev = store.rev();
    }
  }
  if (!data_dir.empty() && store.rev() != snap_rev) write_snapshot(store, data_dir);
  for (auto& kv : conns) close(kv.first);
  close(lfd);
  close(ep);
  return 0;
}
