// hic_huffman.cuh -- host-side Huffman code construction, an exact replay of the reference.
//
// reference hiccup/huffman.py:60-79 builds the tree with Python's heapq over nodes that compare by
// frequency ONLY (huffman.py:249-250), starting from the leaves in first-occurrence order
// (utils.group_by, utils.py:83-96).  Which of two equal-frequency nodes is popped first is decided
// by heapq's sift order, and that decides the codes, so heapq's algorithm (CPython
// Lib/heapq.py: heapify / heappop / heappush with _siftup / _siftdown) is restated operation for
// operation.  First popped = left child = bit '1', second = right = bit '0' (huffman.py:119-129).
// A single leaf gets the code "1" (huffman.py:66-67, 181-182).
#pragma once
#include <stdint.h>
#include <vector>

namespace hic {

struct HuffCode {
    uint64_t bits;     // the code, right aligned, first bit of the code string = most significant
    uint32_t len;
};

class HeapqHuffman {
  public:
    // freqs: leaf frequencies in first-occurrence order.  Returns false if a code exceeds max_len.
    bool build(const uint32_t* freqs, uint32_t n, std::vector<HuffCode>& out, uint32_t max_len = 58) {
        out.assign(n, HuffCode{0, 0});
        if (n == 0) return true;
        if (n == 1) {
            out[0] = HuffCode{1, 1};
            return true;
        }
        freq_.resize(2 * (size_t)n);
        left_.assign(2 * (size_t)n, -1);
        right_.assign(2 * (size_t)n, -1);
        heap_.resize(n);
        for (uint32_t i = 0; i < n; ++i) {
            freq_[i] = freqs[i];
            heap_[i] = (int)i;
        }
        for (int i = (int)(n / 2) - 1; i >= 0; --i) siftup(i);          // heapq.heapify
        int next = (int)n;
        while (heap_.size() > 1) {
            const int l = pop();
            const int r = pop();
            freq_[next] = freq_[l] + freq_[r];
            left_[next] = l;
            right_[next] = r;
            heap_.push_back(next);                                       // heapq.heappush
            siftdown(0, (int)heap_.size() - 1);
            ++next;
        }
        const int root = heap_[0];
        // iterative DFS
        stack_.clear();
        stack_.push_back({root, HuffCode{0, 0}});
        while (!stack_.empty()) {
            auto [node, code] = stack_.back();
            stack_.pop_back();
            if (left_[node] < 0) {
                out[node] = code;
                continue;
            }
            if (code.len + 1 > max_len) return false;
            stack_.push_back({left_[node], HuffCode{(code.bits << 1) | 1, code.len + 1}});
            stack_.push_back({right_[node], HuffCode{(code.bits << 1), code.len + 1}});
        }
        return true;
    }

  private:
    bool lt(int a, int b) const { return freq_[a] < freq_[b]; }

    // heapq._siftdown(heap, startpos, pos)
    void siftdown(int startpos, int pos) {
        const int newitem = heap_[pos];
        while (pos > startpos) {
            const int parentpos = (pos - 1) >> 1;
            const int parent = heap_[parentpos];
            if (lt(newitem, parent)) {
                heap_[pos] = parent;
                pos = parentpos;
                continue;
            }
            break;
        }
        heap_[pos] = newitem;
    }

    // heapq._siftup(heap, pos)
    void siftup(int pos) {
        const int endpos = (int)heap_.size();
        const int startpos = pos;
        const int newitem = heap_[pos];
        int childpos = 2 * pos + 1;
        while (childpos < endpos) {
            const int rightpos = childpos + 1;
            if (rightpos < endpos && !lt(heap_[childpos], heap_[rightpos])) childpos = rightpos;
            heap_[pos] = heap_[childpos];
            pos = childpos;
            childpos = 2 * pos + 1;
        }
        heap_[pos] = newitem;
        siftdown(startpos, pos);
    }

    // heapq.heappop
    int pop() {
        const int lastelt = heap_.back();
        heap_.pop_back();
        if (!heap_.empty()) {
            const int ret = heap_[0];
            heap_[0] = lastelt;
            siftup(0);
            return ret;
        }
        return lastelt;
    }

    struct Frame {
        int node;
        HuffCode code;
    };
    std::vector<uint64_t> freq_;
    std::vector<int> left_, right_, heap_;
    std::vector<Frame> stack_;
};

}  // namespace hic
