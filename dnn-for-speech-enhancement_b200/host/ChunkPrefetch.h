// ChunkPrefetcher — double-buffered chunk prefetch for BPtrain's training loop.
//
// The reference reads a chunk, trains on it, reads the next (BPtrain.cc:45-54, Interface::Readchunk): the GPU idles
// while the host reads and the host idles while the GPU is fed.  Here bp_train / bp_train_raw return as soon as the H2D
// copies of the chunk are complete and its bunches are queued, so the GPU works on chunk i while the host reads chunk
// i+1 — but read -> wait for the upload -> queue ~1500 kernel launches still ran back to back on the one host thread.
// This class runs exactly the serial loop's Interface::Readchunk / ReadchunkRaw calls, in exactly the serial order (so
// the lrand48 stream and every row / table are unchanged), on one reader thread that stays one chunk ahead of the
// consumer in a second set of (page-locked) buffers.
//
// Hand-over protocol: next() releases the slot handed out before (the caller's bp_train* call has returned, i.e. the
// H2D copies out of that slot are complete — bp_gpu.h) and blocks until the following chunk is ready.  The reader may
// fill slot i%2 once chunk i-2 has been released.  During the lifetime of the object the calling thread must not touch
// the Interface's reader state (lrand48, sample_sent / sample_frame_in_sent, the Pfile handles, the chunk buffers).
#pragma once
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <utility>
#include <vector>

class ChunkPrefetcher {
 public:
  using ReadFn = std::function<int(int chunk, int slot)>;  // reads `chunk` into buffer set `slot` (0|1) -> samples

  ChunkPrefetcher(std::vector<int> order, ReadFn read) : order_(std::move(order)), read_(std::move(read)) {
    worker_ = std::thread([this] { run(); });
  }
  ~ChunkPrefetcher() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    worker_.join();
  }
  ChunkPrefetcher(const ChunkPrefetcher&) = delete;
  ChunkPrefetcher& operator=(const ChunkPrefetcher&) = delete;

  // Samples of the next chunk of `order`, *slot = the buffer set it is in (valid until the following call);
  // -1 after the last chunk.
  int next(int* slot) {
    std::unique_lock<std::mutex> lk(mu_);
    released_ = taken_;  // everything handed out so far may be overwritten
    cv_.notify_all();
    if (taken_ >= static_cast<long>(order_.size())) return -1;
    cv_.wait(lk, [&] { return ready_ > taken_; });
    const long i = taken_++;
    *slot = static_cast<int>(i & 1);
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
      const int n = read_(order_[i], static_cast<int>(i & 1));
      {
        std::lock_guard<std::mutex> lk(mu_);
        samples_[i & 1] = n;
        ready_ = i + 1;
      }
      cv_.notify_all();
    }
  }

  std::vector<int> order_;
  ReadFn read_;
  int samples_[2] = {0, 0};
  std::mutex mu_;
  std::condition_variable cv_;
  long ready_ = 0;     // chunks read so far
  long taken_ = 0;     // chunks handed to the consumer
  long released_ = 0;  // the consumer is done with every chunk of index < released_
  bool stop_ = false;
  std::thread worker_;
};
