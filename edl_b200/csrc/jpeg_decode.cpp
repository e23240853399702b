// Batched JPEG decode on the GPU (nvJPEG) feeding the fused crop / resize / normalise kernel (augment.cu).
//
// Reference: DALI's mixed-device decoder of the ImageNet examples (example/distill/resnet/dali.py:60-75,
// example/collective/resnet50/dali.py) -- file bytes go to the device, pixels never come back to the host.  At the
// measured 6.7k img/s per GPU (52k img/s per node) a CPU decoder needs ~100 cores per node; the decode belongs on
// the GPU (hybrid Huffman-on-CPU / IDCT-on-GPU backend by default, the NVJPG engine with backend="hardware").
//
// nvJPEG is loaded with dlopen at first use, so `_C.so` has no link-time dependency on it: a box without
// libnvjpeg.so.12 keeps every other op and the CPU (OpenCV) loader; only JpegDecoder() raises.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <dlfcn.h>
#include <nvjpeg.h>
#include <pybind11/stl.h>
#include <torch/extension.h>

#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "kernels.h"

namespace {

namespace py = pybind11;
using torch::Tensor;

struct NvjpegApi {
  decltype(&nvjpegCreateEx) CreateEx = nullptr;
  decltype(&nvjpegDestroy) Destroy = nullptr;
  decltype(&nvjpegJpegStateCreate) StateCreate = nullptr;
  decltype(&nvjpegJpegStateDestroy) StateDestroy = nullptr;
  decltype(&nvjpegGetImageInfo) GetImageInfo = nullptr;
  decltype(&nvjpegDecode) Decode = nullptr;
  decltype(&nvjpegDecodeBatchedInitialize) BatchedInit = nullptr;
  decltype(&nvjpegDecodeBatched) Batched = nullptr;
  std::string error;
};

template <typename F>
bool load_sym(void* lib, const char* name, F& out, std::string& err) {
  out = reinterpret_cast<F>(dlsym(lib, name));
  if (out == nullptr) err = std::string("nvjpeg symbol missing: ") + name;
  return out != nullptr;
}

const NvjpegApi& api() {
  static NvjpegApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    void* lib = nullptr;
    for (const char* n : {"libnvjpeg.so.12", "/usr/local/cuda/lib64/libnvjpeg.so.12", "libnvjpeg.so"}) {
      lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
      if (lib != nullptr) break;
    }
    if (lib == nullptr) {
      a.error = "libnvjpeg.so.12 not found (dlopen)";
      return;
    }
    load_sym(lib, "nvjpegCreateEx", a.CreateEx, a.error) && load_sym(lib, "nvjpegDestroy", a.Destroy, a.error) &&
        load_sym(lib, "nvjpegJpegStateCreate", a.StateCreate, a.error) &&
        load_sym(lib, "nvjpegJpegStateDestroy", a.StateDestroy, a.error) &&
        load_sym(lib, "nvjpegGetImageInfo", a.GetImageInfo, a.error) &&
        load_sym(lib, "nvjpegDecode", a.Decode, a.error) &&
        load_sym(lib, "nvjpegDecodeBatchedInitialize", a.BatchedInit, a.error) &&
        load_sym(lib, "nvjpegDecodeBatched", a.Batched, a.error);
  });
  return a;
}

void check(nvjpegStatus_t st, const char* what) {
  if (st != NVJPEG_STATUS_SUCCESS)
    throw std::runtime_error(std::string("nvjpeg: ") + what + " failed with status " + std::to_string((int)st));
}

class JpegDecoder {
 public:
  JpegDecoder(int device, const std::string& backend, int cpu_threads) : device_(device), cpu_threads_(cpu_threads) {
    const NvjpegApi& a = api();
    if (!a.error.empty()) throw std::runtime_error(a.error);
    c10::cuda::CUDAGuard g(device_);
    nvjpegBackend_t be = NVJPEG_BACKEND_DEFAULT;
    if (backend == "hardware") be = NVJPEG_BACKEND_HARDWARE;
    else if (backend == "gpu_hybrid") be = NVJPEG_BACKEND_GPU_HYBRID;
    else if (backend == "hybrid") be = NVJPEG_BACKEND_HYBRID;
    else if (backend != "default") throw std::runtime_error("unknown nvjpeg backend " + backend);
    check(a.CreateEx(be, nullptr, nullptr, NVJPEG_FLAGS_DEFAULT, &handle_), "nvjpegCreateEx");
    check(a.StateCreate(handle_, &state_), "nvjpegJpegStateCreate");
    backend_ = backend;
  }
  ~JpegDecoder() {
    const NvjpegApi& a = api();
    if (state_ != nullptr) a.StateDestroy(state_);
    if (handle_ != nullptr) a.Destroy(handle_);
  }
  JpegDecoder(const JpegDecoder&) = delete;
  JpegDecoder& operator=(const JpegDecoder&) = delete;

  // [(height, width, components)] from the JPEG headers (host only, no GPU work)
  std::vector<std::tuple<int, int, int>> image_info(const std::vector<py::bytes>& blobs) {
    std::vector<std::tuple<int, int, int>> out;
    out.reserve(blobs.size());
    for (const py::bytes& b : blobs) {
      char* p;
      Py_ssize_t n;
      if (PyBytes_AsStringAndSize(b.ptr(), &p, &n) != 0) throw py::error_already_set();
      int comps = 0, w[NVJPEG_MAX_COMPONENT], h[NVJPEG_MAX_COMPONENT];
      nvjpegChromaSubsampling_t ss;
      check(api().GetImageInfo(handle_, reinterpret_cast<const unsigned char*>(p), (size_t)n, &comps, &ss, w, h),
            "nvjpegGetImageInfo");
      out.emplace_back(h[0], w[0], comps);
    }
    return out;
  }

  // Decode blobs[i] as interleaved RGB into pool[offsets[i] : offsets[i] + heights[i] * pitches[i]] on the current
  // stream.  The bitstreams are referenced until the NEXT decode() call (the copy to the device may be asynchronous).
  void decode(std::vector<py::bytes> blobs, Tensor& pool, const std::vector<int64_t>& offsets,
              const std::vector<int64_t>& pitches, const std::vector<int64_t>& heights) {
    const size_t n = blobs.size();
    TORCH_CHECK(pool.is_cuda() && pool.scalar_type() == at::kByte && pool.is_contiguous() && pool.dim() == 1);
    TORCH_CHECK(pool.get_device() == device_, "pool lives on another device than the decoder");
    TORCH_CHECK(offsets.size() == n && pitches.size() == n && heights.size() == n);
    std::vector<const unsigned char*> data(n);
    std::vector<size_t> lens(n);
    std::vector<nvjpegImage_t> dst(n);
    uint8_t* base = pool.data_ptr<uint8_t>();
    for (size_t i = 0; i < n; ++i) {
      char* p;
      Py_ssize_t len;
      if (PyBytes_AsStringAndSize(blobs[i].ptr(), &p, &len) != 0) throw py::error_already_set();
      data[i] = reinterpret_cast<const unsigned char*>(p);
      lens[i] = (size_t)len;
      TORCH_CHECK(offsets[i] >= 0 && offsets[i] + heights[i] * pitches[i] <= pool.numel(), "image ", i,
                  " does not fit the pool");
      for (int c = 0; c < NVJPEG_MAX_COMPONENT; ++c) {
        dst[i].channel[c] = nullptr;
        dst[i].pitch[c] = 0;
      }
      dst[i].channel[0] = base + offsets[i];
      dst[i].pitch[0] = (size_t)pitches[i];
    }
    c10::cuda::CUDAGuard g(device_);
    cudaStream_t stream = at::cuda::getCurrentCUDAStream().stream();
    held_ = std::move(blobs);                       // keeps the bitstreams alive; the previous batch is released here
    {
      py::gil_scoped_release nogil;                 // Huffman decoding of the hybrid backends runs on this thread
      const NvjpegApi& a = api();
      if (n == 1) {
        check(a.Decode(handle_, state_, data[0], lens[0], NVJPEG_OUTPUT_RGBI, &dst[0], stream), "nvjpegDecode");
      } else if (n > 1) {
        if ((int)n != batch_) {
          check(a.BatchedInit(handle_, state_, (int)n, cpu_threads_, NVJPEG_OUTPUT_RGBI),
                "nvjpegDecodeBatchedInitialize");
          batch_ = (int)n;
        }
        nvjpegStatus_t st = a.Batched(handle_, state_, data.data(), lens.data(), dst.data(), stream);
        if (st != NVJPEG_STATUS_SUCCESS) {
          batch_ = -1;                              // a failed batch must be re-initialised
          check(st, "nvjpegDecodeBatched");
        }
      }
    }
  }

  std::string backend() const { return backend_; }

 private:
  int device_;
  int cpu_threads_;
  int batch_ = -1;
  std::string backend_;
  nvjpegHandle_t handle_ = nullptr;
  nvjpegJpegState_t state_ = nullptr;
  std::vector<py::bytes> held_;
};

void crop_resize_normalize(const Tensor& pool, const Tensor& items, Tensor& y, std::vector<double> mean,
                           std::vector<double> stdv) {
  TORCH_CHECK(pool.is_cuda() && pool.scalar_type() == at::kByte && pool.is_contiguous());
  TORCH_CHECK(items.is_cuda() && items.scalar_type() == at::kInt && items.is_contiguous() && items.dim() == 2 &&
              items.size(1) == 8, "items: int32 [N, 8] = offset_lo, offset_hi, pitch, y, x, ch, cw, flip");
  TORCH_CHECK(y.is_cuda() && y.scalar_type() == at::kBFloat16 && y.is_contiguous() && y.dim() == 4 &&
              y.size(0) == items.size(0) && y.size(1) == y.size(2) && y.size(3) == 3, "y: bf16 [N, S, S, 3]");
  static_assert(sizeof(edl::AugmentItem) == 32, "AugmentItem must be 8 x int32");
  const float m[3] = {(float)mean[0], (float)mean[1], (float)mean[2]};
  const float sd[3] = {(float)stdv[0], (float)stdv[1], (float)stdv[2]};
  c10::cuda::CUDAGuard g(pool.device());
  edl::crop_resize_normalize(pool.data_ptr<uint8_t>(), reinterpret_cast<const edl::AugmentItem*>(items.data_ptr<int>()),
                             y.data_ptr(), (int)y.size(0), (int)y.size(1), m, sd,
                             at::cuda::getCurrentCUDAStream().stream());
}

bool nvjpeg_available() { return api().error.empty(); }

}  // namespace

void register_jpeg_bindings(pybind11::module_& m) {
  m.def("nvjpeg_available", &nvjpeg_available, "libnvjpeg could be loaded (says nothing about a GPU being present)");
  m.def("crop_resize_normalize", &crop_resize_normalize);
  py::class_<JpegDecoder>(m, "JpegDecoder")
      .def(py::init<int, const std::string&, int>(), py::arg("device"), py::arg("backend") = "default",
           py::arg("cpu_threads") = 4)
      .def("image_info", &JpegDecoder::image_info)
      .def("decode", &JpegDecoder::decode)
      .def_property_readonly("backend", &JpegDecoder::backend);
}
