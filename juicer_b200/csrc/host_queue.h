// host_queue.h — layout of the shared utterance queue (see host_queue.cpp)
#pragma once
#include <atomic>
#include <string>

struct JgpuQueueShared {
    alignas(64) std::atomic<long long> next;        // position in the (longest-first) utterance order
    alignas(64) std::atomic<long long> generation;  // bumped by every create / reset (diagnostics)
};

struct jgpu_queue {
    JgpuQueueShared* shared;
    std::string name;
    bool owner;
};
