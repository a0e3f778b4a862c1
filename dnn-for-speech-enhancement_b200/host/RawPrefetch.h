// RawPrefetcher — double-buffered chunk prefetch for the device-side reader.
//
// The reference reads a chunk, trains on it, reads the next (BPtrain.cc:45-54, Interface::Readchunk): the GPU idles
// while the host reads and the host idles while the GPU is fed.  With reader=gpu the per-chunk host work is the two
// record blocks (page cache -> page-locked buffer) and the sample table; this class runs exactly those
// Interface::ReadchunkRaw calls, in exactly the serial order (so the lrand48 stream and every table are unchanged), on
// one reader thread that stays one chunk ahead of the consumer: while the caller uploads chunk i and queues its
// bunches, chunk i+1 is being read into the other slot.
//
// Hand-over protocol: next() releases the slot handed out before (the caller's bp_train_raw has returned, i.e. the
// H2D copies of that slot are complete — bp_gpu.h) and blocks until the following chunk is ready.  The reader may fill
// slot i%2 once chunk i-2 has been released.  During the lifetime of the object the calling thread must not touch the
// Interface's reader state (lrand48, sample_sent / sample_frame_in_sent, the Pfile handles).
#pragma once
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "Interface.h"

class RawPrefetcher {
 public:
  RawPrefetcher(Interface* io, const std::vector<int>& order) : io_(io), order_(order) {
    worker_ = std::thread([this] { run(); });
  }
  ~RawPrefetcher() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      released_ = static_cast<long>(order_.size());
    }
    cv_.notify_all();
    worker_.join();
    for (RawChunk& s : slot_) io_->free_raw(&s);
  }
  RawPrefetcher(const RawPrefetcher&) = delete;
  RawPrefetcher& operator=(const RawPrefetcher&) = delete;

  // Samples of the next chunk (in `order`), *rc = its slot; valid until the following call.  -1 after the last chunk.
  int next(RawChunk** rc) {
    std::unique_lock<std::mutex> lk(mu_);
    released_ = taken_;  // everything handed out so far may be overwritten
    cv_.notify_all();
    if (taken_ >= static_cast<long>(order_.size())) return -1;
    cv_.wait(lk, [&] { return ready_ > taken_; });
    const long i = taken_++;
    *rc = &slot_[i & 1];
    return samples_[i & 1];
  }

 private:
  void run() {
    for (long i = 0; i < static_cast<long>(order_.size()); ++i) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return stop_ || released_ >= i - 1; });  // chunk i-2 (same slot) has been released
        if (stop_) return;
      }
      const int n = io_->ReadchunkRaw(order_[i], &slot_[i & 1]);
      {
        std::lock_guard<std::mutex> lk(mu_);
        samples_[i & 1] = n;
        ready_ = i + 1;
      }
      cv_.notify_all();
    }
  }

  Interface* io_;
  std::vector<int> order_;
  RawChunk slot_[2];
  int samples_[2] = {0, 0};
  std::mutex mu_;
  std::condition_variable cv_;
  long ready_ = 0;     // chunks read so far
  long taken_ = 0;     // chunks handed to the consumer
  long released_ = 0;  // chunks the consumer is done with: all with index < released_
  bool stop_ = false;
  std::thread worker_;
};
