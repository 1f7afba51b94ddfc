// Minimal stand-in for the oneTBB API surface the SQUANDER hot-path translation units use.
// TEST INFRASTRUCTURE ONLY (oracle build): lets the reference sources under /root/reference
// compile unmodified in an image that has no oneTBB. Not part of the product.
//
// parallel_for / parallel_invoke are backed by OpenMP (outermost level only; nested calls run
// serially on the calling thread, which is also what a saturated TBB arena converges to).
// Set SQREF_SERIAL=1 in the environment (read once) to force everything serial.
#pragma once
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <vector>
#include <omp.h>

inline void* scalable_aligned_malloc(size_t size, size_t align) {
    void* p = nullptr;
    if (size == 0) size = align;
    // TBB's scalable allocator hands out whole cache lines. The reference relies on that slack: the Hilbert-Schmidt
    // correction-2 path writes a 4th element into a Matrix(1,3) (Optimization_Interface.cpp:719-726 vs :1380), which
    // lands in the padding under TBB but would corrupt the glibc heap with an exact-size allocation.
    size = (size + align - 1) / align * align;
    if (posix_memalign(&p, align < sizeof(void*) ? sizeof(void*) : align, size) != 0) return nullptr;
    return p;
}
inline void scalable_aligned_free(void* p) { free(p); }
inline void* scalable_aligned_realloc(void* p, size_t size, size_t align) {
    // no usable-size query in the shim: callers in the reference only grow small arrays
    void* q = scalable_aligned_malloc(size, align);
    if (p && q) { memcpy(q, p, size); free(p); }
    return q;
}
inline void* scalable_malloc(size_t s) { return malloc(s); }
inline void scalable_free(void* p) { free(p); }

namespace tbb {

inline bool shim_serial() {
    static const bool v = [] { const char* e = getenv("SQREF_SERIAL"); return e && e[0] == '1'; }();
    return v;
}

template <typename T>
class blocked_range {
public:
    typedef T const_iterator;
    blocked_range(T b, T e, size_t g = 1) : b_(b), e_(e), g_(g ? g : 1) {}
    T begin() const { return b_; }
    T end() const { return e_; }
    size_t grainsize() const { return g_; }
    size_t size() const { return size_t(e_ - b_); }
    bool empty() const { return !(b_ < e_); }
private:
    T b_, e_;
    size_t g_;
};

class affinity_partitioner {};
class auto_partitioner {};
class simple_partitioner {};
class static_partitioner {};

template <typename T, typename F>
void parallel_for(const blocked_range<T>& r, const F& f) {
    if (r.empty()) return;
    const long long n = (long long)r.size();
    const long long g = (long long)r.grainsize();
    const long long chunks = (n + g - 1) / g;
    if (shim_serial() || omp_in_parallel() || chunks <= 1) {
        f(r);
        return;
    }
    const T b = r.begin();
    const T e = r.end();
#pragma omp parallel for schedule(dynamic, 1)
    for (long long c = 0; c < chunks; ++c) {
        T lo = b + (T)(c * g);
        T hi = (c + 1 == chunks) ? e : (T)(lo + (T)g);
        f(blocked_range<T>(lo, hi, (size_t)g));
    }
}
template <typename T, typename F, typename P>
void parallel_for(const blocked_range<T>& r, const F& f, P&) { parallel_for(r, f); }

template <typename I, typename F>
void parallel_for(I first, I last, I step, const F& f) {
    for (I i = first; i < last; i += step) f(i);
}
template <typename I, typename F>
void parallel_for(I first, I last, const F& f) {
    for (I i = first; i < last; ++i) f(i);
}

// parallel_invoke runs its tasks one after the other: every call site on the hot path pairs one cheap task with one
// that opens a wide parallel_for (Gates_block::apply_to_combined, Gates_block.cpp:1344-1364), and OpenMP -- unlike
// TBB -- would serialise that inner loop if the invoke itself were a 2-thread parallel region.
template <typename F0, typename F1>
void parallel_invoke(const F0& f0, const F1& f1) {
    f0();
    f1();
}
template <typename F0, typename F1, typename F2>
void parallel_invoke(const F0& f0, const F1& f1, const F2& f2) { f0(); f1(); f2(); }

class tick_count {
public:
    class interval_t {
    public:
        interval_t(double s = 0) : s_(s) {}
        double seconds() const { return s_; }
    private:
        double s_;
    };
    static tick_count now() { tick_count t; t.t_ = std::chrono::steady_clock::now(); return t; }
    friend interval_t operator-(const tick_count& a, const tick_count& b) {
        return interval_t(std::chrono::duration<double>(a.t_ - b.t_).count());
    }
private:
    std::chrono::steady_clock::time_point t_;
};

// One slot per OpenMP thread id (outermost team). Slots are created lazily under a lock.
template <typename T>
class enumerable_thread_specific {
public:
    enumerable_thread_specific() : init_([] { return T(); }) {}
    template <typename F, typename = decltype(std::declval<F>()())>
    explicit enumerable_thread_specific(F f) : init_(f) {}
    explicit enumerable_thread_specific(const T& v) : init_([v] { return v; }) {}
    T& local() {
        const int id = slot_id();
        std::lock_guard<std::mutex> lk(m_);
        if ((int)slots_.size() <= id) slots_.resize(id + 1);
        if (!slots_[id]) slots_[id].reset(new T(init_()));
        return *slots_[id];
    }
    void clear() { std::lock_guard<std::mutex> lk(m_); slots_.clear(); }
    template <typename F>
    void combine_each(F f) { for (auto& s : slots_) if (s) f(*s); }
private:
    static int slot_id() {
        // level-1 thread id; nested regions are serial so the ancestor id is stable
        return omp_get_level() == 0 ? 0 : omp_get_ancestor_thread_num(1);
    }
    std::function<T()> init_;
    std::vector<std::unique_ptr<T>> slots_;
    std::mutex m_;
};

template <typename T>
class combinable {
public:
    combinable() : ets_() {}
    template <typename F>
    explicit combinable(F f) : ets_(f) {}
    T& local() { return ets_.local(); }
    template <typename F>
    void combine_each(F f) { ets_.combine_each(f); }
    template <typename F>
    T combine(F f) {
        bool first = true; T acc = T();
        ets_.combine_each([&](T& v) { if (first) { acc = v; first = false; } else acc = f(acc, v); });
        return acc;
    }
    void clear() { ets_.clear(); }
private:
    enumerable_thread_specific<T> ets_;
};

class spin_mutex {
public:
    class scoped_lock {
    public:
        scoped_lock() : m_(nullptr) {}
        explicit scoped_lock(spin_mutex& m) : m_(&m) { m_->m_.lock(); }
        ~scoped_lock() { if (m_) m_->m_.unlock(); }
        void acquire(spin_mutex& m) { m_ = &m; m_->m_.lock(); }
        void release() { if (m_) { m_->m_.unlock(); m_ = nullptr; } }
    private:
        spin_mutex* m_;
    };
    void lock() { m_.lock(); }
    void unlock() { m_.unlock(); }
private:
    std::mutex m_;
};

class queuing_mutex {
public:
    class scoped_lock {
    public:
        scoped_lock() : m_(nullptr) {}
        explicit scoped_lock(queuing_mutex& m) : m_(&m) { m_->m_.lock(); }
        ~scoped_lock() { if (m_) m_->m_.unlock(); }
        void acquire(queuing_mutex& m) { m_ = &m; m_->m_.lock(); }
        void release() { if (m_) { m_->m_.unlock(); m_ = nullptr; } }
    private:
        queuing_mutex* m_;
    };
private:
    std::mutex m_;
};

class queuing_rw_mutex {
public:
    class scoped_lock {
    public:
        scoped_lock() : m_(nullptr) {}
        explicit scoped_lock(queuing_rw_mutex& m, bool /*write*/ = true) : m_(&m) { m_->m_.lock(); }
        ~scoped_lock() { if (m_) m_->m_.unlock(); }
        void acquire(queuing_rw_mutex& m, bool /*write*/ = true) { m_ = &m; m_->m_.lock(); }
        void release() { if (m_) { m_->m_.unlock(); m_ = nullptr; } }
    private:
        queuing_rw_mutex* m_;
    };
private:
    std::mutex m_;
};

template <typename T>
class cache_aligned_allocator : public std::allocator<T> {
public:
    template <typename U> struct rebind { typedef cache_aligned_allocator<U> other; };
    cache_aligned_allocator() {}
    template <typename U> cache_aligned_allocator(const cache_aligned_allocator<U>&) {}
};
template <typename T>
class scalable_allocator : public std::allocator<T> {
public:
    template <typename U> struct rebind { typedef scalable_allocator<U> other; };
    scalable_allocator() {}
    template <typename U> scalable_allocator(const scalable_allocator<U>&) {}
};

class task_group {
public:
    template <typename F> void run(const F& f) { f(); }
    template <typename F> void run_and_wait(const F& f) { f(); }
    void wait() {}
};

class task_arena {
public:
    explicit task_arena(int = 0) {}
    template <typename F> void execute(const F& f) { f(); }
};

}  // namespace tbb
