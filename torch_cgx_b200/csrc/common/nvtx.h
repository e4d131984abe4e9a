// NVTX ranges for Nsight Systems / ncu --nvtx timelines (tracing; the reference has
// only the c10d profiling title, /root/reference/src/ProcessGroupCGX.cc:375-399).
// nvtx3 is header-only: without an attached tool every call is a load + a not-taken branch.
#pragma once
#include <nvtx3/nvToolsExt.h>

#include <cstdint>
#include <cstdio>

namespace cgx {

// RAII range in the "cgx" domain; the payload (bytes, bucket index, ...) shows up in the tool tip.
class NvtxRange {
 public:
  NvtxRange(const char* name, uint64_t payload = 0) {
    nvtxEventAttributes_t a = {};
    a.version = NVTX_VERSION;
    a.size = NVTX_EVENT_ATTRIB_STRUCT_SIZE;
    a.colorType = NVTX_COLOR_ARGB;
    a.color = 0xFF2E8B57u;
    a.payloadType = NVTX_PAYLOAD_TYPE_UNSIGNED_INT64;
    a.payload.ullValue = payload;
    a.messageType = NVTX_MESSAGE_TYPE_ASCII;
    a.message.ascii = name;
    nvtxDomainRangePushEx(domain(), &a);
  }
  ~NvtxRange() { nvtxDomainRangePop(domain()); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;

 private:
  static nvtxDomainHandle_t domain() {
    static nvtxDomainHandle_t d = nvtxDomainCreateA("cgx");
    return d;
  }
};

}  // namespace cgx
