#include "fd_channel.h"

#include <poll.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <atomic>
#include <cerrno>
#include <cstring>
#include <random>
#include <stdexcept>

namespace cgx {
namespace {

// abstract namespace: sun_path[0] == 0, no file system entry, vanishes with the socket
socklen_t make_addr(const std::string& name, sockaddr_un* addr) {
  std::memset(addr, 0, sizeof(*addr));
  addr->sun_family = AF_UNIX;
  const size_t n = std::min(name.size(), sizeof(addr->sun_path) - 2);
  std::memcpy(addr->sun_path + 1, name.data(), n);
  return (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + n);
}

[[noreturn]] void fail(const std::string& what) {
  throw std::runtime_error("cgx: fd channel: " + what + ": " + std::strerror(errno));
}

struct Payload {
  int32_t kind;
  int32_t src;
};

}  // namespace

FdChannel::FdChannel(const std::string& hint) {
  sock_ = ::socket(AF_UNIX, SOCK_DGRAM | SOCK_CLOEXEC, 0);
  if (sock_ < 0) fail("socket");
  static std::atomic<uint64_t> counter{0};
  std::random_device rd;
  for (int attempt = 0; attempt < 16; ++attempt) {
    name_ = "cgx-" + hint + "-" + std::to_string((long)getpid()) + "-" + std::to_string(counter++) + "-" +
            std::to_string(((uint64_t)rd() << 32) | rd());
    sockaddr_un addr;
    const socklen_t len = make_addr(name_, &addr);
    if (::bind(sock_, reinterpret_cast<sockaddr*>(&addr), len) == 0) return;
    if (errno != EADDRINUSE) break;
  }
  const int e = errno;
  ::close(sock_);
  sock_ = -1;
  errno = e;
  fail("bind");
}

FdChannel::~FdChannel() {
  if (sock_ >= 0) ::close(sock_);
}

bool FdChannel::try_send(const std::string& peer_name, int fd, int32_t kind, int32_t src) const {
  sockaddr_un addr;
  const socklen_t len = make_addr(peer_name, &addr);
  Payload pl{kind, src};
  iovec iov{&pl, sizeof(pl)};
  alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
  std::memset(ctrl, 0, sizeof(ctrl));
  msghdr msg;
  std::memset(&msg, 0, sizeof(msg));
  msg.msg_name = &addr;
  msg.msg_namelen = len;
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  msg.msg_control = ctrl;
  msg.msg_controllen = sizeof(ctrl);
  cmsghdr* c = CMSG_FIRSTHDR(&msg);
  c->cmsg_level = SOL_SOCKET;
  c->cmsg_type = SCM_RIGHTS;
  c->cmsg_len = CMSG_LEN(sizeof(int));
  std::memcpy(CMSG_DATA(c), &fd, sizeof(int));
  if (::sendmsg(sock_, &msg, MSG_DONTWAIT) == (ssize_t)sizeof(pl)) return true;
  if (errno == EAGAIN || errno == EWOULDBLOCK || errno == ENOBUFS || errno == EINTR) return false;
  fail("sendmsg to " + peer_name);
}

bool FdChannel::try_recv(FdMessage* out_msg, int timeout_ms) const {
  pollfd pfd{sock_, POLLIN, 0};
  const int pr = ::poll(&pfd, 1, timeout_ms);
  if (pr == 0) return false;
  if (pr < 0) {
    if (errno == EINTR) return false;
    fail("poll");
  }
  Payload pl{0, -1};
  iovec iov{&pl, sizeof(pl)};
  alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
  msghdr msg;
  std::memset(&msg, 0, sizeof(msg));
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  msg.msg_control = ctrl;
  msg.msg_controllen = sizeof(ctrl);
  const ssize_t n = ::recvmsg(sock_, &msg, MSG_CMSG_CLOEXEC);
  if (n != (ssize_t)sizeof(pl)) fail("recvmsg");
  FdMessage out;
  out.kind = pl.kind;
  out.src = pl.src;
  for (cmsghdr* c = CMSG_FIRSTHDR(&msg); c != nullptr; c = CMSG_NXTHDR(&msg, c)) {
    if (c->cmsg_level == SOL_SOCKET && c->cmsg_type == SCM_RIGHTS) {
      std::memcpy(&out.fd, CMSG_DATA(c), sizeof(int));
      break;
    }
  }
  if (out.fd < 0) {
    errno = EPROTO;
    fail("message without a descriptor");
  }
  *out_msg = out;
  return true;
}

}  // namespace cgx
