// thickness_wrapper.cpp -- C++ mirror of org.bonej.wrapperPlugins.ThicknessWrapper and the helpers it
// uses (see include/thickness_wrapper.hpp for the reference citations).  Host code only; the
// numerical work is one call into the C-ABI (bjt_thickness_u8 / bjt_thickness_both_u8).
#include "../../include/thickness_wrapper.hpp"
#include "../../include/bonej_thickness.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <stdexcept>

namespace bonej {

// ---- java.util.Random, as ElementUtil's static RNG = new Random(42) (ElementUtil.java:63) ---------------
namespace {
struct JavaRandom {
    uint64_t seed;
    explicit JavaRandom(uint64_t s) : seed((s ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1)) {}
    int32_t next(int bits) {
        seed = (seed * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
        return (int32_t)(seed >> (48 - bits));
    }
    double nextDouble() { return (double)(((int64_t)next(26) << 27) + next(27)) * (1.0 / (double)(1LL << 53)); }
};
JavaRandom g_rng(42);
std::mutex g_rng_lock;
}  // namespace

long Dataset::size() const {
    long n = 1;
    for (const Axis &a : axes) n *= a.size;
    return n;
}

// ---- SharedTable ------------------------------------------------------------------------------------------
std::mutex SharedTable::lock_;
SharedTable::Table SharedTable::table_;

void SharedTable::add(const std::string &label, const std::string &header, double value) {
    if (label.empty()) throw std::invalid_argument("Label cannot be empty");
    if (header.empty()) throw std::invalid_argument("Header cannot be empty");
    std::lock_guard<std::mutex> lk(lock_);
    Table &t = table_;
    size_t col = std::find(t.headers.begin(), t.headers.end(), header) - t.headers.begin();
    if (col == t.headers.size()) {              // appendEmptyColumn
        t.headers.push_back(header);
        t.columns.emplace_back(t.row_labels.size(), std::numeric_limits<double>::quiet_NaN());
        t.filled.emplace_back(t.row_labels.size(), false);
    }
    // insertIntoNextFreeRow: bottom-up, first row with this label whose cell is empty
    for (size_t i = t.row_labels.size(); i-- > 0;) {
        if (t.row_labels[i] == label && !t.filled[col][i]) {
            t.columns[col][i] = value;
            t.filled[col][i] = true;
            return;
        }
    }
    t.row_labels.push_back(label);
    for (size_t c = 0; c < t.headers.size(); ++c) {
        t.columns[c].push_back(std::numeric_limits<double>::quiet_NaN());
        t.filled[c].push_back(false);
    }
    t.columns[col].back() = value;
    t.filled[col].back() = true;
}

SharedTable::Table SharedTable::getTable() {
    std::lock_guard<std::mutex> lk(lock_);
    return table_;
}

bool SharedTable::hasData() {
    std::lock_guard<std::mutex> lk(lock_);
    for (const auto &c : table_.filled)
        for (bool f : c) if (f) return true;
    return false;
}

void SharedTable::reset() {
    std::lock_guard<std::mutex> lk(lock_);
    table_ = Table();
}

// ---- ElementUtil / AxisUtils / ResultUtils ---------------------------------------------------------------
bool isIJ1Binary(const Dataset &ds, long n) {
    if (ds.bits_per_pixel != 8 || !ds.data) return false;
    const long total = ds.size();
    if (total <= n) {
        for (long i = 0; i < total; ++i)
            if (ds.data[i] != 0 && ds.data[i] != 255) return false;
        return true;
    }
    std::lock_guard<std::mutex> lk(g_rng_lock);
    for (long i = 0; i < n; ++i) {
        // the reference turns the flat index into per-axis positions; with first-axis-fastest data
        // that is the same byte
        const long idx = (long)(g_rng.nextDouble() * (double)total);
        if (ds.data[idx] != 0 && ds.data[idx] != 255) return false;
    }
    return true;
}

static bool isSpatial(AxisType t) { return t == AxisType::X || t == AxisType::Y || t == AxisType::Z; }

bool has3SpatialDimensions(const Dataset &ds) {
    int n = 0;
    for (const Axis &a : ds.axes) n += isSpatial(a.type);
    return n == 3;
}
bool hasChannelDimensions(const Dataset &ds) {
    for (const Axis &a : ds.axes) if (a.type == AxisType::CHANNEL) return true;
    return false;
}
bool hasTimeDimensions(const Dataset &ds) {
    for (const Axis &a : ds.axes) if (a.type == AxisType::TIME) return true;
    return false;
}

// unit conversion of the spatial axes: the reference asks a UnitService; here units must be equal, or
// all empty (uncalibrated), or metric lengths that convert by a power of ten
static bool unitFactor(const std::string &u, double *f) {
    static const struct { const char *u; double f; } T[] = {{"m", 1.0}, {"cm", 1e-2}, {"mm", 1e-3}, {"um", 1e-6},
                                                            {"\xC2\xB5m", 1e-6}, {"micron", 1e-6}, {"nm", 1e-9}};
    for (const auto &e : T) if (u == e.u) { *f = e.f; return true; }
    return false;
}

double getAnisotropy(const Dataset &ds) {
    std::vector<double> scales;
    std::string common;
    bool first = true, convertible = true;
    double f0 = 1.0;
    for (const Axis &a : ds.axes) {
        if (!isSpatial(a.type)) continue;
        double s = a.scale;
        if (first) { common = a.unit; first = false; if (!unitFactor(common, &f0)) f0 = 0.0; }
        else if (a.unit != common) {
            double f;
            if (f0 > 0.0 && unitFactor(a.unit, &f)) s = s * f / f0;
            else convertible = false;
        }
        scales.push_back(s);
    }
    if (scales.empty()) throw std::invalid_argument("Isotropy cannot be determined: no spatial axes");
    if (!convertible)
        throw std::invalid_argument("Isotropy cannot be determined: units of spatial calibrations are inconvertible");
    if (scales.size() < 2) return 0.0;
    std::sort(scales.begin(), scales.end());
    if (scales.front() <= 0.0) return std::numeric_limits<double>::infinity();
    return scales.back() / scales.front() - 1.0;
}

bool isSpatialCalibrationsIsotropic(const Dataset &ds, double tolerance) {
    if (tolerance < 0.0) throw std::invalid_argument("Tolerance cannot be negative");
    if (std::isnan(tolerance)) throw std::invalid_argument("Tolerance cannot be NaN");
    return !(getAnisotropy(ds) - tolerance > 1e-12);
}

std::string getUnitHeader(const Dataset &ds) {
    std::string unit;
    for (const Axis &a : ds.axes) if (a.type == AxisType::X) unit = a.unit;
    if (unit.empty()) unit = "pixel";             // ij.measure.Calibration.getUnit() default
    return "(" + unit + ")";
}

// ---- the command ------------------------------------------------------------------------------------------
void ThicknessWrapper::validateImage() {
    if (!inputDataset) { cancel(messages::NO_IMAGE_OPEN); return; }
    if (!has3SpatialDimensions(*inputDataset)) { cancel(messages::NOT_3D_IMAGE); return; }
    if (hasChannelDimensions(*inputDataset)) {
        cancel(std::string(messages::HAS_CHANNEL_DIMENSIONS) + ". Please split the channels.");
        return;
    }
    if (hasTimeDimensions(*inputDataset)) {
        cancel(std::string(messages::HAS_TIME_DIMENSIONS) + ". Please split the hyperstack.");
        return;
    }
    if (!isIJ1Binary(*inputDataset, 1000000)) { cancel(messages::NOT_BINARY); return; }
    if (!anisotropyWarned_) {
        if (!isSpatialCalibrationsIsotropic(*inputDataset, 0.01)) {
            char pct[64];
            snprintf(pct, sizeof pct, "(%.1f %%)", getAnisotropy(*inputDataset) * 100.0);
            if (headless) {
                log.push_back(std::string("WARN The image is anisotropic ") + pct +
                              ". Continuing anyway, but results may be unreliable");
            } else if (!anisotropyDialogAnswerOk) {
                cancel("");                          // cancel(null)
            }
        }
        anisotropyWarned_ = true;
    }
}

void ThicknessWrapper::run() {
    // getMapOptions
    std::vector<bool> options;
    if (mapChoice == "Trabecular thickness") options = {true};
    else if (mapChoice == "Trabecular separation") options = {false};
    else if (mapChoice == "Both") options = {true, false};
    else throw std::invalid_argument("Unexpected map choice");

    const Dataset &ds = *inputDataset;
    long w = 1, h = 1, d = 1;
    double pixelWidth = 1.0;
    for (const Axis &a : ds.axes) {
        if (a.type == AxisType::X) { w = a.size; pixelWidth = a.scale; }
        else if (a.type == AxisType::Y) h = a.size;
        else if (a.type == AxisType::Z) d = a.size;
    }
    const size_t N = (size_t)w * h * d;

    bjt_ctx *ctx = nullptr;
    if (bjt_create(&ctx, device) != BJT_OK)
        throw std::runtime_error(std::string("Thickness: ") + bjt_last_error(nullptr));

    std::unique_ptr<ThicknessMap> maps[2];   // [0] foreground (Tb.Th), [1] background (Tb.Sp)
    bjt_stats stats[2];
    bool have[2] = {false, false};
    auto newMap = [&](bool foreground) {
        std::unique_ptr<ThicknessMap> m(new ThicknessMap());
        m->name = ds.name + (foreground ? "_Tb.Th" : "_Tb.Sp");
        m->w = w; m->h = h; m->d = d; m->axes = ds.axes;
        if (showMaps) m->data.resize(N);
        return m;
    };
    int rc = BJT_OK;
    log.push_back("STATUS Thickness: creating thickness map");
    if (options.size() == 2) {
        maps[0] = newMap(true); maps[1] = newMap(false);
        rc = bjt_thickness_both_u8(ctx, ds.data, (int)w, (int)h, (int)d, maskArtefacts ? 1 : 0, 128, pixelWidth,
                                   showMaps ? maps[0]->data.data() : nullptr, showMaps ? maps[1]->data.data() : nullptr,
                                   &stats[0], &stats[1], nullptr, nullptr);
        have[0] = have[1] = (rc == BJT_OK);
    } else {
        const bool fg = options[0];
        const int k = fg ? 0 : 1;
        maps[k] = newMap(fg);
        rc = bjt_thickness_u8(ctx, ds.data, (int)w, (int)h, (int)d, fg ? 0 : 1, maskArtefacts ? 1 : 0, 128, pixelWidth,
                              showMaps ? maps[k]->data.data() : nullptr, &stats[k], nullptr);
        have[k] = (rc == BJT_OK);
    }
    std::string err = rc == BJT_OK ? "" : bjt_last_error(ctx);
    bjt_destroy(ctx);
    // processImage returning null (not a 0/255 image) makes addMapResults return silently (:158-161);
    // any other failure is an exception out of the module -- there is no CPU path to fall back to
    if (rc != BJT_OK && rc != BJT_E_NOT_BINARY) throw std::runtime_error("Thickness: " + err);

    log.push_back("STATUS Thickness: calculating results");
    const std::string unitHeader = getUnitHeader(ds);
    for (int k = 0; k < 2; ++k) {
        if (!have[k]) continue;
        const std::string prefix = k == 0 ? "Tb.Th" : "Tb.Sp";
        double mean = stats[k].mean, sd = stats[k].sd, mx = stats[k].max;
        if (stats[k].count == 0) mean = sd = mx = std::numeric_limits<double>::quiet_NaN();
        SharedTable::add(ds.name, prefix + " Mean " + unitHeader, mean);
        SharedTable::add(ds.name, prefix + " Std Dev " + unitHeader, sd);
        SharedTable::add(ds.name, prefix + " Max " + unitHeader, mx);
    }
    resultsTable = SharedTable::getTable();
    if (showMaps) {
        for (int k = 0; k < 2; ++k) {
            if (!have[k]) continue;
            maps[k]->channel_min = 0.0;
            maps[k]->channel_max = stats[k].max;
            maps[k]->color_table = "FIRE";
            (k == 0 ? trabecularMap : separationMap) = std::move(maps[k]);
        }
    }
}

void ThicknessWrapper::execute() {
    canceled_ = false;
    cancelReason_.clear();
    validateImage();
    if (!canceled_) run();
}

}  // namespace bonej

// ---- C entry points ---------------------------------------------------------------------------------------
extern "C" {

int bjw_thickness_run(const bjw_dataset *ds, const char *map_choice, int show_maps, int mask_artefacts, int headless,
                      int dialog_ok, int device, bjw_result *out) {
    if (!out) return -1;
    memset(out, 0, sizeof *out);
    bonej::Dataset d;
    if (ds) {
        d.name = ds->name ? ds->name : "";
        d.bits_per_pixel = ds->bits_per_pixel;
        d.data = ds->data;
        for (int i = 0; i < ds->n_axes; ++i) {
            bonej::Axis a;
            a.type = (bonej::AxisType)ds->axis_types[i];
            a.size = ds->axis_sizes[i];
            a.scale = ds->axis_scales ? ds->axis_scales[i] : 1.0;
            a.unit = (ds->axis_units && ds->axis_units[i]) ? ds->axis_units[i] : "";
            d.axes.push_back(a);
        }
    }
    bonej::ThicknessWrapper cmd;
    cmd.inputDataset = ds ? &d : nullptr;
    if (map_choice) cmd.mapChoice = map_choice;
    cmd.showMaps = show_maps != 0;
    cmd.maskArtefacts = mask_artefacts != 0;
    cmd.headless = headless != 0;
    cmd.anisotropyDialogAnswerOk = dialog_ok != 0;
    cmd.device = device;
    try {
        cmd.execute();
    } catch (const std::invalid_argument &e) {
        out->exception = 1;
        snprintf(out->cancel_reason, sizeof out->cancel_reason, "%s", e.what());
        return 0;
    } catch (const std::exception &e) {
        out->exception = 2;
        snprintf(out->cancel_reason, sizeof out->cancel_reason, "%s", e.what());
        return -1;
    }
    out->canceled = cmd.isCanceled();
    if (cmd.isCanceled()) snprintf(out->cancel_reason, sizeof out->cancel_reason, "%s", cmd.getCancelReason().c_str());
    std::string lg;
    for (const std::string &l : cmd.log) lg += l + "\n";
    snprintf(out->log, sizeof out->log, "%s", lg.c_str());
    auto copy = [](const std::unique_ptr<bonej::ThicknessMap> &m, float **dst, double *mx) {
        if (!m) return;
        *dst = (float *)malloc(m->data.size() * sizeof(float) + 4);
        memcpy(*dst, m->data.data(), m->data.size() * sizeof(float));
        *mx = m->channel_max;
    };
    copy(cmd.trabecularMap, &out->trabecular_map, &out->trabecular_max);
    copy(cmd.separationMap, &out->separation_map, &out->separation_max);
    if (cmd.trabecularMap) snprintf(out->color_table, sizeof out->color_table, "%s", cmd.trabecularMap->color_table.c_str());
    else if (cmd.separationMap) snprintf(out->color_table, sizeof out->color_table, "%s", cmd.separationMap->color_table.c_str());
    const bonej::SharedTable::Table &t = cmd.resultsTable;
    out->n_columns = (int)std::min<size_t>(t.headers.size(), 16);
    out->n_rows = (int)std::min<size_t>(t.row_labels.size(), 16);
    for (int c = 0; c < out->n_columns; ++c) {
        snprintf(out->headers[c], sizeof out->headers[c], "%s", t.headers[c].c_str());
        for (int r = 0; r < out->n_rows; ++r)
            out->values[c][r] = t.filled[c][r] ? t.columns[c][r] : std::numeric_limits<double>::quiet_NaN();
    }
    for (int r = 0; r < out->n_rows; ++r) snprintf(out->row_labels[r], sizeof out->row_labels[r], "%s", t.row_labels[r].c_str());
    return 0;
}

void bjw_result_free(bjw_result *r) {
    if (!r) return;
    free(r->trabecular_map);
    free(r->separation_map);
    r->trabecular_map = r->separation_map = nullptr;
}

void bjw_shared_table_reset(void) { bonej::SharedTable::reset(); }

}  // extern "C"
