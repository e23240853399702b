// Symmetric memory for one elastic stage, bootstrapped through the job's rendezvous store.
//
// The reference re-creates NCCL communicators on every stage change by broadcasting an ncclUniqueId over TCP
// among the restarted trainers (python/edl/utils/train_process.py:37-41,55).  Here a stage's communication
// fabric is nothing but memory: every rank creates ONE physical slab with the CUDA virtual-memory-management
// API (cuMemCreate), exports it as a POSIX file descriptor, and maps every peer's slab (cuMemImport... +
// cuMemMap) plus -- when the NVSwitch supports it -- a multicast object bound to all slabs (cuMulticastCreate /
// cuMulticastBindMem), whose alias address is what the `multimem.*` instructions of csrc/allreduce.cu target.
// No NCCL communicator exists on this path.  What travels through the key-value store is only the NAME of each
// rank's handle server (an abstract unix socket that passes the descriptors with SCM_RIGHTS); the store itself
// is the job's KV store or whatever `torch.distributed.Store` the launcher provided (parallel/symm.py).
//
// The driver API is reached through dlopen("libcuda.so.1"): `_C.so` has no link-time dependency on it, so the
// extension still imports (and build() still passes) on a box without a GPU driver.
#include <cuda.h>
#include <dlfcn.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <torch/extension.h>
#include <unistd.h>

#include <atomic>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Driver {
  void* lib = nullptr;
  std::string error;
#define EDL_DRV(name) decltype(&::name) name = nullptr;
  EDL_DRV(cuInit)
  EDL_DRV(cuGetErrorString)
  EDL_DRV(cuDeviceGet)
  EDL_DRV(cuDeviceGetAttribute)
  EDL_DRV(cuDevicePrimaryCtxRetain)
  EDL_DRV(cuCtxSetCurrent)
  EDL_DRV(cuMemGetAllocationGranularity)
  EDL_DRV(cuMemCreate)
  EDL_DRV(cuMemRelease)
  EDL_DRV(cuMemAddressReserve)
  EDL_DRV(cuMemAddressFree)
  EDL_DRV(cuMemMap)
  EDL_DRV(cuMemUnmap)
  EDL_DRV(cuMemSetAccess)
  EDL_DRV(cuMemExportToShareableHandle)
  EDL_DRV(cuMemImportFromShareableHandle)
  EDL_DRV(cuMulticastCreate)
  EDL_DRV(cuMulticastAddDevice)
  EDL_DRV(cuMulticastBindMem)
  EDL_DRV(cuMulticastUnbind)
  EDL_DRV(cuMulticastGetGranularity)
#undef EDL_DRV
};

Driver& drv() {
  static Driver d;
  static std::once_flag once;
  std::call_once(once, [] {
    d.lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (d.lib == nullptr) {
      d.error = "libcuda.so.1 not found (no GPU driver on this box)";
      return;
    }
    auto sym = [&](const char* versioned, const char* plain) -> void* {
      void* p = dlsym(d.lib, versioned);
      return p != nullptr ? p : dlsym(d.lib, plain);
    };
#define EDL_LOAD(name)                                                          \
  d.name = reinterpret_cast<decltype(d.name)>(sym(#name "_v2", #name));         \
  if (d.name == nullptr && d.error.empty()) d.error = "libcuda.so.1 lacks " #name;
    EDL_LOAD(cuInit)
    EDL_LOAD(cuGetErrorString)
    EDL_LOAD(cuDeviceGet)
    EDL_LOAD(cuDeviceGetAttribute)
    EDL_LOAD(cuDevicePrimaryCtxRetain)
    EDL_LOAD(cuCtxSetCurrent)
    EDL_LOAD(cuMemGetAllocationGranularity)
    EDL_LOAD(cuMemCreate)
    EDL_LOAD(cuMemRelease)
    EDL_LOAD(cuMemAddressReserve)
    EDL_LOAD(cuMemAddressFree)
    EDL_LOAD(cuMemMap)
    EDL_LOAD(cuMemUnmap)
    EDL_LOAD(cuMemSetAccess)
    EDL_LOAD(cuMemExportToShareableHandle)
    EDL_LOAD(cuMemImportFromShareableHandle)
    EDL_LOAD(cuMulticastCreate)
    EDL_LOAD(cuMulticastAddDevice)
    EDL_LOAD(cuMulticastBindMem)
    EDL_LOAD(cuMulticastUnbind)
    EDL_LOAD(cuMulticastGetGranularity)
#undef EDL_LOAD
  });
  return d;
}

void check(CUresult r, const char* what) {
  if (r == CUDA_SUCCESS) return;
  const char* s = nullptr;
  if (drv().cuGetErrorString != nullptr) drv().cuGetErrorString(r, &s);
  TORCH_CHECK(false, "vmm: ", what, " failed: CUresult=", (int)r, " (", s ? s : "?", ")");
}

void bind_context(int device) {
  auto& d = drv();
  TORCH_CHECK(d.error.empty(), "vmm: ", d.error);
  check(d.cuInit(0), "cuInit");
  CUdevice dev;
  check(d.cuDeviceGet(&dev, device), "cuDeviceGet");
  CUcontext ctx;
  check(d.cuDevicePrimaryCtxRetain(&ctx, dev), "cuDevicePrimaryCtxRetain");   // the context torch uses
  check(d.cuCtxSetCurrent(ctx), "cuCtxSetCurrent");
}

size_t round_up(size_t n, size_t m) { return (n + m - 1) / m * m; }

// ---------------------------------------------------------------------------------------------------------
// Handle server: passes this rank's file descriptors to peers over an abstract unix socket (SCM_RIGHTS).
// Request = one byte (index of the descriptor), reply = one byte status + the descriptor as ancillary data.
class FdServer {
 public:
  explicit FdServer(const std::string& name) : name_(name) {
    sock_ = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
    TORCH_CHECK(sock_ >= 0, "vmm: socket(): ", strerror(errno));
    sockaddr_un addr{};
    addr.sun_family = AF_UNIX;
    TORCH_CHECK(name.size() + 1 < sizeof(addr.sun_path), "vmm: socket name too long");
    memcpy(addr.sun_path + 1, name.data(), name.size());          // leading NUL = abstract namespace
    socklen_t len = (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + name.size());
    if (bind(sock_, reinterpret_cast<sockaddr*>(&addr), len) != 0 || listen(sock_, 64) != 0) {
      int e = errno;
      close(sock_);
      TORCH_CHECK(false, "vmm: bind/listen on @", name, ": ", strerror(e));
    }
    thread_ = std::thread([this] { serve(); });
  }
  ~FdServer() { stop(); }
  void set_fd(int index, int fd) {
    std::lock_guard<std::mutex> g(mu_);
    if ((int)fds_.size() <= index) fds_.resize(index + 1, -1);
    fds_[index] = fd;
  }
  void stop() {
    if (stopped_.exchange(true)) return;
    shutdown(sock_, SHUT_RDWR);
    close(sock_);
    if (thread_.joinable()) thread_.join();
  }
  const std::string& name() const { return name_; }

 private:
  void serve() {
    while (!stopped_.load()) {
      int c = accept4(sock_, nullptr, nullptr, SOCK_CLOEXEC);
      if (c < 0) {
        if (stopped_.load() || (errno != EINTR && errno != ECONNABORTED)) return;
        continue;
      }
      unsigned char idx = 0;
      if (recv(c, &idx, 1, MSG_WAITALL) == 1) {
        int fd = -1;
        {
          std::lock_guard<std::mutex> g(mu_);
          if (idx < fds_.size()) fd = fds_[idx];
        }
        char status = fd >= 0 ? 1 : 0;
        iovec io{&status, 1};
        msghdr msg{};
        msg.msg_iov = &io;
        msg.msg_iovlen = 1;
        alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
        if (fd >= 0) {
          memset(ctrl, 0, sizeof(ctrl));
          msg.msg_control = ctrl;
          msg.msg_controllen = sizeof(ctrl);
          cmsghdr* cm = CMSG_FIRSTHDR(&msg);
          cm->cmsg_level = SOL_SOCKET;
          cm->cmsg_type = SCM_RIGHTS;
          cm->cmsg_len = CMSG_LEN(sizeof(int));
          memcpy(CMSG_DATA(cm), &fd, sizeof(int));
        }
        sendmsg(c, &msg, MSG_NOSIGNAL);
      }
      close(c);
    }
  }
  std::string name_;
  int sock_ = -1;
  std::thread thread_;
  std::mutex mu_;
  std::vector<int> fds_;
  std::atomic<bool> stopped_{false};
};

int fetch_fd(const std::string& name, int index) {
  int s = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
  TORCH_CHECK(s >= 0, "vmm: socket(): ", strerror(errno));
  sockaddr_un addr{};
  addr.sun_family = AF_UNIX;
  TORCH_CHECK(name.size() + 1 < sizeof(addr.sun_path), "vmm: socket name too long");
  memcpy(addr.sun_path + 1, name.data(), name.size());
  socklen_t len = (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + name.size());
  if (connect(s, reinterpret_cast<sockaddr*>(&addr), len) != 0) {
    int e = errno;
    close(s);
    TORCH_CHECK(false, "vmm: connect to @", name, ": ", strerror(e));
  }
  unsigned char idx = (unsigned char)index;
  char status = 0;
  int fd = -1;
  bool ok = send(s, &idx, 1, MSG_NOSIGNAL) == 1;
  if (ok) {
    iovec io{&status, 1};
    msghdr msg{};
    msg.msg_iov = &io;
    msg.msg_iovlen = 1;
    alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
    memset(ctrl, 0, sizeof(ctrl));
    msg.msg_control = ctrl;
    msg.msg_controllen = sizeof(ctrl);
    ok = recvmsg(s, &msg, MSG_CMSG_CLOEXEC) == 1 && status == 1;
    if (ok) {
      cmsghdr* cm = CMSG_FIRSTHDR(&msg);
      ok = cm != nullptr && cm->cmsg_level == SOL_SOCKET && cm->cmsg_type == SCM_RIGHTS;
      if (ok) memcpy(&fd, CMSG_DATA(cm), sizeof(int));
    }
  }
  close(s);
  TORCH_CHECK(ok && fd >= 0, "vmm: peer @", name, " did not hand out descriptor ", index);
  return fd;
}

// ---------------------------------------------------------------------------------------------------------
// One rank's view of a stage's symmetric slab.  Python (parallel/symm.py) drives the steps in lock-step with
// the peers through the rendezvous store:
//   1. SymmSlab(device, nbytes, world, rank, name)  allocate + map locally, start the handle server
//   2. map_peer(r, peer_server_name)                 for every other rank
//   3. rank 0: mc_create(); others: mc_import(rank0 name);  everybody: mc_add_device()     [store barrier]
//   4. mc_bind()                                                                             [store barrier]
class SymmSlab : public std::enable_shared_from_this<SymmSlab> {
 public:
  SymmSlab(int device, int64_t nbytes, int world, int rank, const std::string& server_name)
      : device_(device), world_(world), rank_(rank) {
    bind_context(device);
    auto& d = drv();
    prop_ = CUmemAllocationProp{};
    prop_.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop_.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop_.location.id = device;
    prop_.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t gran = 0;
    check(d.cuMemGetAllocationGranularity(&gran, &prop_, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED), "cuMemGetAllocationGranularity");
    mc_supported_ = false;
    if (world > 1) {
      int v = 0;
      CUdevice dev;
      check(d.cuDeviceGet(&dev, device), "cuDeviceGet");
      if (d.cuDeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) == CUDA_SUCCESS && v != 0) {
        CUmulticastObjectProp mp{};
        mp.numDevices = (unsigned)world;
        mp.size = round_up((size_t)nbytes, gran);
        mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
        size_t mg = 0;
        if (d.cuMulticastGetGranularity(&mg, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && mg != 0) {
          gran = std::max(gran, mg);
          mc_supported_ = true;
        }
      }
    }
    size_ = round_up((size_t)nbytes, gran);
    gran_ = gran;
    check(d.cuMemCreate(&mem_, size_, &prop_, 0), "cuMemCreate");
    have_mem_ = true;
    ptrs_.assign(world, 0);
    peer_mem_.assign(world, 0);
    ptrs_[rank] = map_handle(mem_);
    int fd = -1;
    check(d.cuMemExportToShareableHandle(&fd, mem_, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0), "cuMemExportToShareableHandle");
    fds_.push_back(fd);
    server_ = std::make_unique<FdServer>(server_name);
    server_->set_fd(0, fd);
  }

  ~SymmSlab() { release(); }

  void map_peer(int peer, const std::string& peer_server) {
    TORCH_CHECK(peer >= 0 && peer < world_ && peer != rank_ && ptrs_[peer] == 0, "vmm: bad peer ", peer);
    bind_context(device_);
    int fd = fetch_fd(peer_server, 0);
    CUmemGenericAllocationHandle h;
    CUresult r = drv().cuMemImportFromShareableHandle(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
    close(fd);
    check(r, "cuMemImportFromShareableHandle");
    peer_mem_[peer] = h;
    ptrs_[peer] = map_handle(h);
  }

  bool mc_supported() const { return mc_supported_; }

  void mc_create() {
    TORCH_CHECK(mc_supported_ && !have_mc_, "vmm: multicast unavailable");
    bind_context(device_);
    CUmulticastObjectProp mp{};
    mp.numDevices = (unsigned)world_;
    mp.size = size_;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    check(drv().cuMulticastCreate(&mc_, &mp), "cuMulticastCreate");
    have_mc_ = true;
    int fd = -1;
    check(drv().cuMemExportToShareableHandle(&fd, mc_, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0), "export multicast handle");
    fds_.push_back(fd);
    server_->set_fd(1, fd);
  }
  void mc_import(const std::string& root_server) {
    TORCH_CHECK(mc_supported_ && !have_mc_, "vmm: multicast unavailable");
    bind_context(device_);
    int fd = fetch_fd(root_server, 1);
    CUresult r = drv().cuMemImportFromShareableHandle(&mc_, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
    close(fd);
    check(r, "import multicast handle");
    have_mc_ = true;
  }
  void mc_add_device() {
    TORCH_CHECK(have_mc_);
    bind_context(device_);
    CUdevice dev;
    check(drv().cuDeviceGet(&dev, device_), "cuDeviceGet");
    check(drv().cuMulticastAddDevice(mc_, dev), "cuMulticastAddDevice");
  }
  // every rank's device must have been added before the first bind (store barrier in Python)
  void mc_bind() {
    TORCH_CHECK(have_mc_ && mc_ptr_ == 0);
    bind_context(device_);
    check(drv().cuMulticastBindMem(mc_, 0, mem_, 0, size_, 0), "cuMulticastBindMem");
    mc_bound_ = true;
    mc_ptr_ = map_handle(mc_);
  }

  torch::Tensor tensor() {
    auto self = shared_from_this();
    auto opts = torch::TensorOptions().dtype(torch::kUInt8).device(torch::kCUDA, device_);
    return torch::from_blob(reinterpret_cast<void*>(ptrs_[rank_]), {(int64_t)size_}, [self](void*) {}, opts);
  }
  std::vector<int64_t> ptrs() const { return std::vector<int64_t>(ptrs_.begin(), ptrs_.end()); }
  int64_t mc_ptr() const { return (int64_t)mc_ptr_; }
  int64_t size() const { return (int64_t)size_; }
  int64_t granularity() const { return (int64_t)gran_; }
  std::string server_name() const { return server_ ? server_->name() : std::string(); }
  void stop_server() {
    if (server_) server_->stop();
  }

  void release() {
    if (released_) return;
    released_ = true;
    auto& d = drv();
    if (server_) server_->stop();
    for (int fd : fds_) close(fd);
    fds_.clear();
    if (!d.error.empty()) return;
    if (d.cuCtxSetCurrent != nullptr) {
      CUdevice dev;
      CUcontext ctx;
      if (d.cuDeviceGet(&dev, device_) == CUDA_SUCCESS && d.cuDevicePrimaryCtxRetain(&ctx, dev) == CUDA_SUCCESS)
        d.cuCtxSetCurrent(ctx);
    }
    if (mc_ptr_ != 0) {
      d.cuMemUnmap(mc_ptr_, size_);
      d.cuMemAddressFree(mc_ptr_, size_);
      mc_ptr_ = 0;
    }
    if (mc_bound_) {
      CUdevice dev;
      if (d.cuDeviceGet(&dev, device_) == CUDA_SUCCESS) d.cuMulticastUnbind(mc_, dev, 0, size_);
      mc_bound_ = false;
    }
    if (have_mc_) {
      d.cuMemRelease(mc_);
      have_mc_ = false;
    }
    for (int r = 0; r < world_; ++r) {
      if (ptrs_[r] != 0) {
        d.cuMemUnmap(ptrs_[r], size_);
        d.cuMemAddressFree(ptrs_[r], size_);
        ptrs_[r] = 0;
      }
      if (r != rank_ && peer_mem_[r] != 0) {
        d.cuMemRelease(peer_mem_[r]);
        peer_mem_[r] = 0;
      }
    }
    if (have_mem_) {
      d.cuMemRelease(mem_);
      have_mem_ = false;
    }
  }

 private:
  CUdeviceptr map_handle(CUmemGenericAllocationHandle h) {
    auto& d = drv();
    CUdeviceptr p = 0;
    check(d.cuMemAddressReserve(&p, size_, gran_, 0, 0), "cuMemAddressReserve");
    CUresult r = d.cuMemMap(p, size_, 0, h, 0);
    if (r != CUDA_SUCCESS) {
      d.cuMemAddressFree(p, size_);
      check(r, "cuMemMap");
    }
    CUmemAccessDesc acc{};
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = device_;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    r = d.cuMemSetAccess(p, size_, &acc, 1);
    if (r != CUDA_SUCCESS) {
      d.cuMemUnmap(p, size_);
      d.cuMemAddressFree(p, size_);
      check(r, "cuMemSetAccess");
    }
    return p;
  }

  int device_, world_, rank_;
  CUmemAllocationProp prop_{};
  size_t size_ = 0, gran_ = 0;
  CUmemGenericAllocationHandle mem_ = 0, mc_ = 0;
  bool have_mem_ = false, have_mc_ = false, mc_bound_ = false, mc_supported_ = false, released_ = false;
  std::vector<CUdeviceptr> ptrs_;
  std::vector<CUmemGenericAllocationHandle> peer_mem_;
  CUdeviceptr mc_ptr_ = 0;
  std::vector<int> fds_;
  std::unique_ptr<FdServer> server_;
};

// descriptor passing alone (no CUDA): lets the CPU test-suite exercise the handle server
int fd_roundtrip_selftest(const std::string& name) {
  int p[2];
  TORCH_CHECK(pipe(p) == 0);
  FdServer srv(name);
  srv.set_fd(0, p[1]);
  int w = fetch_fd(name, 0);
  const char msg[] = "edl";
  bool ok = write(w, msg, 3) == 3;
  char buf[4] = {0, 0, 0, 0};
  ok = ok && read(p[0], buf, 3) == 3 && memcmp(buf, msg, 3) == 0;
  close(w);
  close(p[0]);
  close(p[1]);
  srv.stop();
  return ok ? 1 : 0;
}

}  // namespace

void register_vmm_bindings(pybind11::module_& m) {
  m.def("vmm_driver_error", [] { return drv().error; }, "empty when libcuda.so.1 and the VMM entry points are available");
  m.def("vmm_fd_selftest", &fd_roundtrip_selftest);
  pybind11::class_<SymmSlab, std::shared_ptr<SymmSlab>>(m, "SymmSlab")
      .def(pybind11::init<int, int64_t, int, int, const std::string&>())
      .def("map_peer", &SymmSlab::map_peer, pybind11::call_guard<pybind11::gil_scoped_release>())
      .def("mc_supported", &SymmSlab::mc_supported)
      .def("mc_create", &SymmSlab::mc_create)
      .def("mc_import", &SymmSlab::mc_import, pybind11::call_guard<pybind11::gil_scoped_release>())
      .def("mc_add_device", &SymmSlab::mc_add_device)
      .def("mc_bind", &SymmSlab::mc_bind)
      .def("tensor", &SymmSlab::tensor)
      .def("ptrs", &SymmSlab::ptrs)
      .def("mc_ptr", &SymmSlab::mc_ptr)
      .def("size", &SymmSlab::size)
      .def("granularity", &SymmSlab::granularity)
      .def("server_name", &SymmSlab::server_name)
      .def("stop_server", &SymmSlab::stop_server)
      .def("release", &SymmSlab::release);
}
