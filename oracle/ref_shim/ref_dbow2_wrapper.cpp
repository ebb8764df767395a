// ref_dbow2_wrapper.cpp — C entry points over the vendored DBoW2 functions of the path (Thirdparty/DBoW2/DBoW2), sliced into
// oracle/_ref/gen/dbow2_*.inc by oracle/build_ref.sh.  TEST INFRASTRUCTURE ONLY.  The class skeletons below carry the data
// members and enumerators the slices use (names and enumerator order as in BowVector.h, FeatureVector.h, FORB.h,
// ScoringObject.h, TemplatedVocabulary.h); every function body comes from the reference.
#include <cfloat>
#include <cmath>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "cv_shim.h"

using namespace std;

namespace DBoW2 {

typedef unsigned int WordId;          // BowVector.h
typedef double WordValue;
typedef unsigned int NodeId;          // FeatureVector.h
enum LNorm { L1, L2 };
enum WeightingType { TF_IDF, TF, IDF, BINARY };
enum ScoringType { L1_NORM, L2_NORM, CHI_SQUARE, KL, BHATTACHARYYA, DOT_PRODUCT };

class BowVector : public std::map<WordId, WordValue> {
public:
    void addWeight(WordId id, WordValue v);
    void addIfNotExist(WordId id, WordValue v);
    void normalize(LNorm norm_type);
};
class FeatureVector : public std::map<NodeId, std::vector<unsigned int>> {
public:
    void addFeature(NodeId id, unsigned int i_feature);
};
#include "dbow2_bowvector.inc"
#include "dbow2_featurevector.inc"

class FORB {
public:
    typedef cv::Mat TDescriptor;
    static const int L;
    static int distance(const TDescriptor& a, const TDescriptor& b);
    static void fromString(TDescriptor& a, const std::string& s);
};
const int FORB::L = 32;               // FORB.cpp:24
#include "dbow2_forb_distance.inc"
#include "dbow2_forb_fromstring.inc"

class GeneralScoring {
public:
    virtual double score(const BowVector& v, const BowVector& w) const = 0;
    virtual bool mustNormalize(LNorm& norm) const = 0;
    virtual ~GeneralScoring() {}
};
class L1Scoring : public GeneralScoring {   // __SCORING_CLASS(L1Scoring, true, L1), ScoringObject.h:48-74
public:
    virtual double score(const BowVector& v, const BowVector& w) const;
    virtual inline bool mustNormalize(LNorm& norm) const { norm = L1; return true; }
};
#include "dbow2_l1_score.inc"

template <class TDescriptor, class F>
class TemplatedVocabulary {
public:
    struct Node {                      // TemplatedVocabulary.h:233-272
        NodeId id;
        WordValue weight;
        vector<NodeId> children;
        NodeId parent;
        TDescriptor descriptor;
        WordId word_id;
        Node() : id(0), weight(0), parent(0), word_id(0) {}
        inline bool isLeaf() const { return children.empty(); }
    };
    int m_k = 0, m_L = 0;
    WeightingType m_weighting = TF_IDF;
    ScoringType m_scoring = L1_NORM;
    GeneralScoring* m_scoring_object = nullptr;
    std::vector<Node> m_nodes;
    std::vector<Node*> m_words;
    ~TemplatedVocabulary() { delete m_scoring_object; }
    inline bool empty() const { return m_words.empty(); }
    void createScoringObject() { delete m_scoring_object; m_scoring_object = new L1Scoring; }   // ORBvoc.txt: scoring 0 = L1_NORM
    bool loadFromTextFile(const std::string& filename);
    void transform(const std::vector<TDescriptor>& features, BowVector& v, FeatureVector& fv, int levelsup) const;
    void transform(const TDescriptor& feature, WordId& word_id, WordValue& weight, NodeId* nid = NULL, int levelsup = 0) const;
};
#include "dbow2_transform_all.inc"
#include "dbow2_transform_one.inc"
#include "dbow2_load_text.inc"

}  // namespace DBoW2

typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> ORBVocabulary;   // include/ORBVocabulary.h

extern "C" {

void* refd_vocab_load_text(const char* path)
{
    ORBVocabulary* v = new ORBVocabulary;
    if (!v->loadFromTextFile(path)) { delete v; return nullptr; }
    return v;
}
void refd_vocab_destroy(void* h) { delete (ORBVocabulary*)h; }
void refd_vocab_info(void* h, int* n_nodes, int* n_words, int* k, int* L, int* weighting)
{
    ORBVocabulary* v = (ORBVocabulary*)h;
    *n_nodes = (int)v->m_nodes.size(); *n_words = (int)v->m_words.size(); *k = v->m_k; *L = v->m_L; *weighting = (int)v->m_weighting;
}

// transform(feature, word_id, weight, &nid, levelsup) per feature
void refd_transform_features(void* h, const uint8_t* desc, int n, int levelsup, uint32_t* word, double* weight, uint32_t* node)
{
    ORBVocabulary* v = (ORBVocabulary*)h;
    for (int i = 0; i < n; ++i) {
        DBoW2::WordId w; DBoW2::WordValue wt; DBoW2::NodeId nid = 0;
        v->transform(cv::Mat(1, 32, CV_8U, (void*)(desc + (size_t)i * 32)), w, wt, &nid, levelsup);
        word[i] = w; weight[i] = wt; node[i] = nid;
    }
}

// transform(features, BowVector, FeatureVector, levelsup): BowVector as (ids, values), FeatureVector as CSR
void refd_transform(void* h, const uint8_t* desc, int n, int levelsup, uint32_t* bow_ids, double* bow_vals, int* n_bow, uint32_t* fv_nodes,
                    int32_t* fv_off, uint32_t* fv_idx, int* n_fv)
{
    ORBVocabulary* v = (ORBVocabulary*)h;
    std::vector<cv::Mat> feats;
    for (int i = 0; i < n; ++i) feats.push_back(cv::Mat(1, 32, CV_8U, (void*)(desc + (size_t)i * 32)));
    DBoW2::BowVector bv; DBoW2::FeatureVector fv;
    v->transform(feats, bv, fv, levelsup);
    int t = 0;
    for (auto& kv : bv) { bow_ids[t] = kv.first; bow_vals[t] = kv.second; ++t; }
    *n_bow = t;
    int f = 0, o = 0;
    fv_off[0] = 0;
    for (auto& kv : fv) {
        fv_nodes[f] = kv.first;
        for (unsigned int i : kv.second) fv_idx[o++] = i;
        fv_off[++f] = o;
    }
    *n_fv = f;
}

double refd_score_l1(const uint32_t* ids_a, const double* vals_a, int na, const uint32_t* ids_b, const double* vals_b, int nb)
{
    DBoW2::BowVector a, b;
    for (int i = 0; i < na; ++i) a[ids_a[i]] = vals_a[i];
    for (int i = 0; i < nb; ++i) b[ids_b[i]] = vals_b[i];
    DBoW2::L1Scoring s;
    return s.score(a, b);
}

}  // extern "C"
