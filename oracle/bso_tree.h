// ORACLE -- TEST INFRASTRUCTURE ONLY. Not part of the product; nothing under baby_shark_b200/ may
// include, link or call this. CPU restatement (C++17, f32, -ffp-contract=off) of the reference's
// sparse VDB-style tree `dynamic_vdb!(f32, par 5, 4, 3)`:
//   src/voxel/volume/mod.rs:8, src/voxel/init.rs:1-43,
//   src/voxel/leaf_node/{mod,tree_node,flood_fill,csg}.rs,
//   src/voxel/internal_node/{mod,tree_node,flood_fill,csg}.rs,
//   src/voxel/root_node/{mod,tree_node,flood_fill,csg}.rs, src/voxel/value/f32.rs:8-31.
// The node classes keep the reference's shape (Root = ordered map of 4096^3 nodes -> 32^3 internal
// -> 16^3 internal -> 8^3 leaf) so that flood-fill / CSG / fast-sweep quirks carry over literally.
// One deliberate difference: the reference's internal node stores `union{Box<child>, tile}` and
// leaves slots uninitialised at allocation; here branch pointer and tile value are separate fields
// and the tile value starts as +0.0, so reads the reference would make from uninitialised or
// dangling-pointer bytes are defined (and documented in DESIGN.md as "undefined in reference").
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>
#include <cfloat>
#include <map>
#include <array>
#include <vector>
#include <memory>
#include <algorithm>
#include <functional>

namespace bso {

typedef int64_t idx_t;
struct Vec3i { idx_t x, y, z; };
inline Vec3i operator+(Vec3i a, Vec3i b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline bool operator==(Vec3i a, Vec3i b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

struct Vec3f { float x, y, z; };
inline bool operator==(const Vec3f& a, const Vec3f& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

enum Sign { Positive = 0, Negative = 1 };

// value/f32.rs:8-31
inline Sign sign_of(float v) { return std::signbit(v) ? Negative : Positive; }
inline void set_sign(float& v, Sign s) { v = std::copysign(v, s == Negative ? -1.0f : 1.0f); }
inline float far_value() { return FLT_MAX; }

// voxel/utils.rs:39-61
inline float partial_min(float a, float b) { return (a < b) ? a : b; }
inline float partial_max(float a, float b) { return (a > b) ? a : b; }

template <int N> struct Bits {
    static const int WORDS = (N + 63) / 64;
    uint64_t w[WORDS];
    Bits() { off_all(); }
    bool at(size_t i) const { return (w[i >> 6] >> (i & 63)) & 1; }
    void on(size_t i) { w[i >> 6] |= (uint64_t(1) << (i & 63)); }
    void off(size_t i) { w[i >> 6] &= ~(uint64_t(1) << (i & 63)); }
    void set(size_t i, bool v) { if (v) on(i); else off(i); }
    void off_all() { std::memset(w, 0, sizeof(w)); }
    void on_all() { for (int i = 0; i < N; ++i) on(i); }
    bool is_empty() const { for (int i = 0; i < WORDS; ++i) if (w[i]) return false; return true; }
    bool is_full() const { for (int i = 0; i < N; ++i) if (!at(i)) return false; return true; }
    // data_structures/bitset.rs:127-140: first set bit in ascending linear offset
    long find_first_on() const { for (int i = 0; i < N; ++i) if (at(i)) return i; return -1; }
    void or_with(const Bits& o) { for (int i = 0; i < WORDS; ++i) w[i] |= o.w[i]; }
};

template <class T> struct Tile { Vec3i origin; size_t size; T value; };

// ------------------------------------------------------------------------------------------------
// Leaf: leaf_node/mod.rs:9-36, tree_node.rs
template <class T, int LOG2>
struct LeafNode {
    static const int BRANCHING = LOG2;
    static const int BRANCHING_TOTAL = LOG2;
    static const int SIZE = 1 << (3 * LOG2);
    static const bool IS_LEAF = true;
    typedef T Value;
    typedef LeafNode Leaf;
    template <class U> using As = LeafNode<U, LOG2>;

    T values[SIZE];
    Bits<SIZE> value_mask;
    Vec3i origin_;

    static size_t resolution() { return size_t(1) << BRANCHING_TOTAL; }
    static size_t offset(const Vec3i& i) {
        const idx_t m = (idx_t(1) << BRANCHING_TOTAL) - 1;
        return size_t(((i.x & m) << (2 * LOG2)) + ((i.y & m) << LOG2) + (i.z & m));
    }
    static LeafNode* empty(Vec3i origin) {
        LeafNode* n = new LeafNode();
        n->origin_ = origin;
        for (int i = 0; i < SIZE; ++i) n->values[i] = T();
        return n;
    }
    const T* at(const Vec3i& i) const { size_t o = offset(i); return value_mask.at(o) ? &values[o] : nullptr; }
    void insert(const Vec3i& i, T v) { size_t o = offset(i); value_mask.on(o); values[o] = v; }
    void remove(const Vec3i& i) { value_mask.off(offset(i)); }
    bool is_empty() const { return value_mask.is_empty(); }
    Vec3i origin() const { return origin_; }
    void fill(T v) { value_mask.on_all(); for (int i = 0; i < SIZE; ++i) values[i] = v; }
    void clear() { value_mask.off_all(); }
    LeafNode* clone() const { LeafNode* n = new LeafNode(*this); return n; }
    template <class U, class F> As<U>* clone_map(const F& f) const {
        As<U>* n = As<U>::empty(origin_);
        for (int i = 0; i < SIZE; ++i) if (value_mask.at(i)) { n->value_mask.on(i); n->values[i] = f(values[i]); }
        return n;
    }
    template <class V> void visit_leafs(V& v) const { v.dense(*this); }
    template <class F> void visit_values_mut(F& f) { for (int i = 0; i < SIZE; ++i) if (value_mask.at(i)) f(values[i]); }
    const Leaf* leaf_at(const Vec3i&) const { return this; }
    void remove_empty_branches() {}
    template <class P> void remove_if(P pred) { for (int i = 0; i < SIZE; ++i) if (value_mask.at(i) && pred(values[i])) value_mask.off(i); }

    // leaf_node/flood_fill.rs:12-70 (T = float only)
    void flood_fill() {
        if (value_mask.is_full()) return;
        long first = value_mask.find_first_on();
        if (first < 0) return;
        Sign i = sign_of(values[first]);
        const int R = 1 << LOG2;
        for (int x = 0; x < R; ++x) {
            int x00 = x << (2 * LOG2);
            if (value_mask.at(x00)) i = sign_of(values[x00]);
            Sign j = i;
            for (int y = 0; y < R; ++y) {
                int xy0 = x00 + (y << LOG2);
                if (value_mask.at(xy0)) j = sign_of(values[xy0]);
                Sign k = j;
                for (int z = 0; z < R; ++z) {
                    int xyz = xy0 + z;
                    if (value_mask.at(xyz)) k = sign_of(values[xyz]);
                    else { values[xyz] = far_value(); set_sign(values[xyz], k); }
                }
            }
        }
    }
    void fill_with_sign(Sign s) {
        for (int i = 0; i < SIZE; ++i) { if (!value_mask.at(i)) values[i] = far_value(); set_sign(values[i], s); }
    }
    Sign first_value_sign() const { return sign_of(values[0]); }
    Sign last_value_sign() const { return sign_of(values[SIZE - 1]); }
    Sign sign_at(const Vec3i& i) const { return sign_of(values[offset(i)]); }

    // leaf_node/csg.rs:17-45 ; `other` is consumed
    void csg_union(LeafNode* o) { for (int i = 0; i < SIZE; ++i) values[i] = partial_min(values[i], o->values[i]); value_mask.or_with(o->value_mask); delete o; }
    void csg_subtract(LeafNode* o) { for (int i = 0; i < SIZE; ++i) values[i] = partial_max(values[i], -o->values[i]); value_mask.or_with(o->value_mask); delete o; }
    void csg_intersect(LeafNode* o) { for (int i = 0; i < SIZE; ++i) values[i] = partial_max(values[i], o->values[i]); value_mask.or_with(o->value_mask); delete o; }
    void flip_signs() { for (int i = 0; i < SIZE; ++i) values[i] = -values[i]; }
    void destroy() { delete this; }
};

// ------------------------------------------------------------------------------------------------
// Internal node: internal_node/mod.rs:16-170, tree_node.rs
template <class T, class TChild, int LOG2>
struct InternalNode {
    static const int BRANCHING = LOG2;
    static const int BRANCHING_TOTAL = LOG2 + TChild::BRANCHING_TOTAL;
    static const int SIZE = 1 << (3 * LOG2);
    static const bool IS_LEAF = false;
    typedef T Value;
    typedef TChild Child;
    typedef typename TChild::Leaf Leaf;
    template <class U> using As = InternalNode<U, typename TChild::template As<U>, LOG2>;

    Vec3i origin_;
    Bits<SIZE> child_mask, value_mask;
    TChild** branch;  // [SIZE]
    T* tile;          // [SIZE]

    static size_t resolution() { return size_t(1) << BRANCHING_TOTAL; }
    static size_t offset(const Vec3i& i) {
        const idx_t m = (idx_t(1) << BRANCHING_TOTAL) - 1;
        const int cs = TChild::BRANCHING_TOTAL;
        return size_t((((i.x & m) >> cs) << (2 * LOG2)) + (((i.y & m) >> cs) << LOG2) + ((i.z & m) >> cs));
    }
    Vec3i offset_to_global_index(size_t off) const {
        idx_t x = off >> (2 * LOG2);
        off &= (size_t(1) << (2 * LOG2)) - 1;
        idx_t y = off >> LOG2;
        idx_t z = off & ((size_t(1) << LOG2) - 1);
        const int cs = TChild::BRANCHING_TOTAL;
        return Vec3i{(x << cs) + origin_.x, (y << cs) + origin_.y, (z << cs) + origin_.z};
    }
    static InternalNode* empty(Vec3i origin) {
        InternalNode* n = new InternalNode();
        n->origin_ = origin;
        n->branch = new TChild*[SIZE];
        n->tile = new T[SIZE];
        for (int i = 0; i < SIZE; ++i) { n->branch[i] = nullptr; n->tile[i] = T(); }
        return n;
    }
    void destroy() {
        for (int i = 0; i < SIZE; ++i) if (child_mask.at(i)) branch[i]->destroy();
        delete[] branch; delete[] tile; delete this;
    }
    TChild* add_branch(size_t off) {
        if (child_mask.at(off)) return branch[off];
        child_mask.on(off); value_mask.off(off);
        branch[off] = TChild::empty(offset_to_global_index(off));
        return branch[off];
    }
    TChild* remove_branch(size_t off) {  // returns ownership
        if (!child_mask.at(off)) return nullptr;
        child_mask.off(off);
        TChild* c = branch[off]; branch[off] = nullptr; return c;
    }
    // remove_child (mod.rs:84-98): drops branch or deactivates tile
    void remove_child(size_t off) {
        if (child_mask.at(off)) { child_mask.off(off); branch[off]->destroy(); branch[off] = nullptr; }
        else if (value_mask.at(off)) value_mask.off(off);
    }
    const T* at(const Vec3i& i) const {
        size_t o = offset(i);
        if (child_mask.at(o)) return branch[o]->at(i);
        if (value_mask.at(o)) return &tile[o];
        return nullptr;
    }
    void insert(const Vec3i& i, T v) {
        size_t o = offset(i);
        if (child_mask.at(o)) { branch[o]->insert(i, v); return; }
        if (value_mask.at(o)) {
            T tv = tile[o];
            if (tv == v) return;
            TChild* b = add_branch(o); b->fill(tv); b->insert(i, v); return;
        }
        add_branch(o)->insert(i, v);
    }
    bool is_empty() const { return child_mask.is_empty() && value_mask.is_empty(); }
    Vec3i origin() const { return origin_; }
    void clear() {
        for (int i = 0; i < SIZE; ++i) if (child_mask.at(i)) { branch[i]->destroy(); branch[i] = nullptr; }
        child_mask.off_all(); value_mask.off_all();
    }
    void fill(T v) { clear(); value_mask.on_all(); for (int i = 0; i < SIZE; ++i) tile[i] = v; }
    InternalNode* clone() const {
        InternalNode* n = empty(origin_);
        n->child_mask = child_mask; n->value_mask = value_mask;
        for (int i = 0; i < SIZE; ++i) { n->tile[i] = tile[i]; if (child_mask.at(i)) n->branch[i] = branch[i]->clone(); }
        return n;
    }
    template <class U, class F> As<U>* clone_map(const F& f) const {
        As<U>* n = As<U>::empty(origin_);
        for (int i = 0; i < SIZE; ++i) {
            if (child_mask.at(i)) { n->child_mask.on(i); n->branch[i] = branch[i]->template clone_map<U>(f); }
            else if (value_mask.at(i)) { n->value_mask.on(i); n->tile[i] = f(tile[i]); }
        }
        return n;
    }
    template <class V> void visit_leafs(V& v) const {
        for (int i = 0; i < SIZE; ++i) {
            if (child_mask.at(i)) branch[i]->visit_leafs(v);
            else if (value_mask.at(i)) v.tile(Tile<T>{offset_to_global_index(i), TChild::resolution(), tile[i]});
        }
    }
    template <class F> void visit_values_mut(F& f) {
        for (int i = 0; i < SIZE; ++i) { if (child_mask.at(i)) branch[i]->visit_values_mut(f); else if (value_mask.at(i)) f(tile[i]); }
    }
    const Leaf* leaf_at(const Vec3i& i) const { size_t o = offset(i); return child_mask.at(o) ? branch[o]->leaf_at(i) : nullptr; }
    Leaf* take_leaf_at(const Vec3i& i) {
        size_t o = offset(i);
        if (!child_mask.at(o)) return nullptr;
        if constexpr (TChild::IS_LEAF) return remove_branch(o);
        else return branch[o]->take_leaf_at(i);
    }
    void insert_leaf_at(Leaf* leaf) {
        Vec3i i = leaf->origin();
        size_t o = offset(i);
        value_mask.off(o);
        if constexpr (TChild::IS_LEAF) {
            TChild* old = remove_branch(o); if (old) old->destroy();
            child_mask.on(o); value_mask.off(o); branch[o] = leaf;
        } else {
            if (child_mask.at(o)) branch[o]->insert_leaf_at(leaf);
            else add_branch(o)->insert_leaf_at(leaf);
        }
    }
    void remove_empty_branches() {
        for (int i = 0; i < SIZE; ++i) if (child_mask.at(i)) {
            branch[i]->remove_empty_branches();
            if (branch[i]->is_empty()) { TChild* c = remove_branch(i); c->destroy(); }
        }
    }
    template <class P> void remove_if(P pred) {
        for (int i = 0; i < SIZE; ++i) {
            if (child_mask.at(i)) { branch[i]->remove_if(pred); if (branch[i]->is_empty()) { TChild* c = remove_branch(i); c->destroy(); } }
            else if (value_mask.at(i)) { if (pred(tile[i])) value_mask.off(i); }
        }
    }

    // internal_node/flood_fill.rs:17-117
    void flood_fill() {
        if (value_mask.is_full()) return;
        for (int o = 0; o < SIZE; ++o) if (child_mask.at(o)) branch[o]->flood_fill();
        long fv = value_mask.find_first_on(), fb = child_mask.find_first_on();
        Sign i;
        if (fv >= 0 && fb >= 0) i = (fv <= fb) ? sign_of(tile[fv]) : branch[fb]->first_value_sign();
        else if (fv >= 0) i = sign_of(tile[fv]);
        else if (fb >= 0) i = branch[fb]->first_value_sign();
        else return;
        const int R = 1 << LOG2;
        for (int x = 0; x < R; ++x) {
            int x00 = x << (2 * LOG2);
            if (child_mask.at(x00)) i = branch[x00]->last_value_sign(); else if (value_mask.at(x00)) i = sign_of(tile[x00]);
            Sign j = i;
            for (int y = 0; y < R; ++y) {
                int xy0 = x00 + (y << LOG2);
                if (child_mask.at(xy0)) j = branch[xy0]->last_value_sign(); else if (value_mask.at(xy0)) j = sign_of(tile[xy0]);
                Sign k = j;
                for (int z = 0; z < R; ++z) {
                    int xyz = xy0 + z;
                    if (child_mask.at(xyz)) k = branch[xyz]->last_value_sign();
                    else if (value_mask.at(xyz)) k = sign_of(tile[xyz]);
                    else { tile[xyz] = far_value(); set_sign(tile[xyz], k); }
                }
            }
        }
    }
    void fill_with_sign(Sign s) {
        if (value_mask.is_full()) return;
        for (int i = 0; i < SIZE; ++i) {
            if (child_mask.at(i)) branch[i]->fill_with_sign(s);
            else if (value_mask.at(i)) set_sign(tile[i], s);
            else { tile[i] = far_value(); set_sign(tile[i], s); }
        }
    }
    Sign first_value_sign() const { return child_mask.at(0) ? branch[0]->first_value_sign() : sign_of(tile[0]); }
    // quirk kept: a branch in the LAST slot reports its FIRST value sign (flood_fill.rs:103-108)
    Sign last_value_sign() const { return child_mask.at(SIZE - 1) ? branch[SIZE - 1]->first_value_sign() : sign_of(tile[SIZE - 1]); }
    Sign sign_at(const Vec3i& i) const { size_t o = offset(i); return child_mask.at(o) ? branch[o]->sign_at(i) : sign_of(tile[o]); }

    // internal_node/csg.rs:20-163
    bool is_inside_tile(size_t o) const { return !child_mask.at(o) && sign_of(tile[o]) == Negative; }
    bool is_outside_tile(size_t o) const { return !child_mask.at(o) && sign_of(tile[o]) == Positive; }
    void take_child(InternalNode* other, size_t o) {
        if (other->child_mask.at(o)) {
            other->child_mask.set(o, child_mask.at(o));
            other->value_mask.set(o, value_mask.at(o));
            child_mask.on(o); value_mask.off(o);
            std::swap(branch[o], other->branch[o]);
            std::swap(tile[o], other->tile[o]);
        }
    }
    void make_child_inside(size_t o) { remove_child(o); value_mask.on(o); tile[o] = -far_value(); }
    void csg_union(InternalNode* other) {
        for (int o = 0; o < SIZE; ++o) {
            if (is_inside_tile(o)) continue;
            if (other->is_inside_tile(o)) { make_child_inside(o); continue; }
            if (is_outside_tile(o)) { take_child(other, o); continue; }
            if (child_mask.at(o) && other->child_mask.at(o)) {
                TChild* ob = other->remove_branch(o);
                branch[o]->csg_union(ob);
                if (branch[o]->is_empty()) make_child_inside(o);
            } else {
                other->remove_child(o);
            }
        }
        other->destroy();
    }
    void csg_subtract(InternalNode* other) {
        for (int o = 0; o < SIZE; ++o) {
            if (is_outside_tile(o) || other->is_outside_tile(o)) continue;
            if (other->is_inside_tile(o)) { remove_child(o); continue; }
            if (is_inside_tile(o)) {
                take_child(other, o);
                if (child_mask.at(o)) branch[o]->flip_signs();
                else if (value_mask.at(o)) tile[o] = -tile[o];
                continue;
            }
            if (child_mask.at(o) && other->child_mask.at(o)) { TChild* ob = other->remove_branch(o); branch[o]->csg_subtract(ob); }
            else other->remove_child(o);
        }
        other->destroy();
    }
    void csg_intersect(InternalNode* other) {
        for (int o = 0; o < SIZE; ++o) {
            if (is_outside_tile(o) || other->is_inside_tile(o)) continue;
            if (other->is_outside_tile(o)) { remove_child(o); continue; }
            if (is_inside_tile(o)) { take_child(other, o); continue; }
            if (child_mask.at(o) && other->child_mask.at(o)) { TChild* ob = other->remove_branch(o); branch[o]->csg_intersect(ob); }
            else other->remove_child(o);
        }
        other->destroy();
    }
    void flip_signs() {
        for (int o = 0; o < SIZE; ++o) { if (child_mask.at(o)) branch[o]->flip_signs(); else if (value_mask.at(o)) tile[o] = -tile[o]; }
    }
};

// ------------------------------------------------------------------------------------------------
// Root: root_node/mod.rs:13-32, tree_node.rs, flood_fill.rs:8-63, csg.rs:9-58
template <class TChild>
struct RootNode {
    typedef typename TChild::Value Value;
    typedef TChild Child;
    typedef typename TChild::Leaf Leaf;
    template <class U> using As = RootNode<typename TChild::template As<U>>;
    typedef std::array<idx_t, 3> Key;  // lexicographic (x,y,z) like RootKey::cmp
    std::map<Key, TChild*> root;

    ~RootNode() { clear(); }
    RootNode() {}
    RootNode(const RootNode&) = delete;
    RootNode& operator=(const RootNode&) = delete;

    static Key root_key(const Vec3i& i) {
        const idx_t m = ~((idx_t(1) << TChild::BRANCHING_TOTAL) - 1);
        return Key{i.x & m, i.y & m, i.z & m};
    }
    static Vec3i kv(const Key& k) { return Vec3i{k[0], k[1], k[2]}; }
    const Value* at(const Vec3i& i) const { auto it = root.find(root_key(i)); return it == root.end() ? nullptr : it->second->at(i); }
    void insert(const Vec3i& i, Value v) {
        Key k = root_key(i);
        auto it = root.find(k);
        if (it == root.end()) it = root.emplace(k, TChild::empty(kv(k))).first;
        it->second->insert(i, v);
    }
    bool is_empty() const { for (auto& kvp : root) if (!kvp.second->is_empty()) return false; return true; }
    void clear() { for (auto& kvp : root) kvp.second->destroy(); root.clear(); }
    RootNode* clone() const { RootNode* r = new RootNode(); for (auto& kvp : root) r->root[kvp.first] = kvp.second->clone(); return r; }
    template <class U, class F> As<U>* clone_map(const F& f) const {
        As<U>* r = new As<U>();
        for (auto& kvp : root) r->root[kvp.first] = kvp.second->template clone_map<U>(f);
        return r;
    }
    template <class V> void visit_leafs(V& v) const { for (auto& kvp : root) kvp.second->visit_leafs(v); }
    template <class F> void visit_values_mut(F& f) { for (auto& kvp : root) kvp.second->visit_values_mut(f); }
    const Leaf* leaf_at(const Vec3i& i) const { auto it = root.find(root_key(i)); return it == root.end() ? nullptr : it->second->leaf_at(i); }
    Leaf* take_leaf_at(const Vec3i& i) { auto it = root.find(root_key(i)); return it == root.end() ? nullptr : it->second->take_leaf_at(i); }
    void insert_leaf_at(Leaf* leaf) {
        Key k = root_key(leaf->origin());
        auto it = root.find(k);
        if (it == root.end()) it = root.emplace(k, TChild::empty(kv(k))).first;
        it->second->insert_leaf_at(leaf);
    }
    void remove_empty_branches() {
        for (auto it = root.begin(); it != root.end();) {
            it->second->remove_empty_branches();
            if (it->second->is_empty()) { it->second->destroy(); it = root.erase(it); } else ++it;
        }
    }
    template <class P> void remove_if(P pred) {
        for (auto it = root.begin(); it != root.end();) {
            it->second->remove_if(pred);
            if (it->second->is_empty()) { it->second->destroy(); it = root.erase(it); } else ++it;
        }
    }
    void flood_fill() {
        if (root.empty()) return;
        for (auto& kvp : root) kvp.second->flood_fill();
        std::vector<Key> origins;
        for (auto& kvp : root) origins.push_back(kvp.first);
        const idx_t res = idx_t(TChild::resolution());
        for (size_t n = 0; n + 1 < origins.size(); ++n) {
            const Key a = origins[n], b = origins[n + 1];
            if (a[0] != b[0] || a[1] != b[1] || b[2] == a[2] + res) continue;
            if (root[a]->last_value_sign() != Negative || root[b]->first_value_sign() != Negative) continue;
            Vec3i t{a[0], a[1], a[2] + res};
            while (t.z < b[2]) {
                Key k = root_key(t);
                TChild* node = TChild::empty(kv(k));
                node->fill_with_sign(Negative);
                auto it = root.find(k);
                if (it != root.end()) { it->second->destroy(); it->second = node; } else root[k] = node;
                t.z += res;
            }
        }
    }
    Sign sign_at(const Vec3i& i) const { auto it = root.find(root_key(i)); return it == root.end() ? Positive : it->second->sign_at(i); }
    void csg_union(RootNode* other) {
        std::vector<Key> keys;
        for (auto& kvp : root) keys.push_back(kvp.first);
        for (auto& kvp : other->root) keys.push_back(kvp.first);
        std::sort(keys.begin(), keys.end());
        keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
        for (auto& k : keys) {
            auto a = root.find(k); auto b = other->root.find(k);
            if (b == other->root.end()) continue;
            TChild* n2 = b->second; other->root.erase(b);
            if (a != root.end()) a->second->csg_union(n2); else root[k] = n2;
        }
        delete other;
    }
    void csg_subtract(RootNode* other) {
        std::vector<Key> keys;
        for (auto& kvp : other->root) keys.push_back(kvp.first);
        for (auto& k : keys) {
            auto a = root.find(k);
            if (a == root.end()) continue;
            auto b = other->root.find(k);
            TChild* n2 = b->second; other->root.erase(b);
            a->second->csg_subtract(n2);
        }
        delete other;
    }
    void csg_intersect(RootNode* other) {
        // root_node/csg.rs:36-51: drop keys not in both, then intersect the rest
        for (auto it = root.begin(); it != root.end();) {
            if (other->root.find(it->first) == other->root.end()) { it->second->destroy(); it = root.erase(it); } else ++it;
        }
        std::vector<Key> keys;
        for (auto& kvp : root) keys.push_back(kvp.first);
        for (auto& k : keys) {
            auto b = other->root.find(k);
            if (b == other->root.end()) continue;
            TChild* n2 = b->second; other->root.erase(b);
            root[k]->csg_intersect(n2);
        }
        delete other;
    }
};

template <class T> using Leaf3 = LeafNode<T, 3>;
template <class T> using Node4 = InternalNode<T, Leaf3<T>, 4>;
template <class T> using Node5 = InternalNode<T, Node4<T>, 5>;
template <class T> using Grid = RootNode<Node5<T>>;
typedef Grid<float> VolumeGrid;

}  // namespace bso
