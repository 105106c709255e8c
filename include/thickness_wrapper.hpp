// thickness_wrapper.hpp -- host-side mirror of the reference's Thickness command, in C++ because no
// JVM / JDK exists in the build image (java/ holds the Java sources a maintainer drops in instead).
//
// It keeps the operator interface of org.bonej.wrapperPlugins.ThicknessWrapper
// (Modern/wrapperPlugins/src/main/java/org/bonej/wrapperPlugins/ThicknessWrapper.java:75-261):
// same parameter names, option strings, validation order, cancel reasons, output nullness and
// results-table columns; only createMap() (:190-196) is replaced, by one call into the C-ABI of
// include/bonej_thickness.h.  Helpers mirror
//   SharedTable.add / getTable / reset   Modern/utilities/.../SharedTable.java:118-176,234-254
//   ElementUtil.isIJ1Binary              Modern/utilities/.../ElementUtil.java:173-228
//   AxisUtils (spatial / channel / time axes, anisotropy)  Modern/utilities/.../AxisUtils.java
//   ResultUtils.getUnitHeader(ImagePlus) Modern/wrapperPlugins/.../wrapperUtils/ResultUtils.java:160-167
//   CommonMessages                       Modern/wrapperPlugins/.../CommonMessages.java:38-52
#pragma once
#include <cstdint>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

struct bjt_ctx;

namespace bonej {

// ---- CommonMessages ---------------------------------------------------------------------------------
namespace messages {
constexpr const char *NOT_3D_IMAGE = "Need a 3D (X, Y, Z) image";
constexpr const char *HAS_TIME_DIMENSIONS = "Image cannot have time axis";
constexpr const char *HAS_CHANNEL_DIMENSIONS = "Image cannot be composite";
constexpr const char *NOT_BINARY = "Need a binary image";
constexpr const char *NO_IMAGE_OPEN = "No image open";
}  // namespace messages

// ---- a net.imagej.Dataset reduced to what the command reads --------------------------------------------
enum class AxisType { X, Y, Z, CHANNEL, TIME, UNKNOWN };

struct Axis {
    AxisType type = AxisType::UNKNOWN;
    long size = 1;
    double scale = 1.0;          // averageScale(0, 1)
    std::string unit;            // "" = uncalibrated
};

struct Dataset {
    std::string name;
    std::vector<Axis> axes;      // first axis fastest, as in imglib2
    int bits_per_pixel = 8;      // 8 = UnsignedByteType; anything else fails isIJ1Binary
    const uint8_t *data = nullptr;

    long size() const;
};

// ---- the map outputs (a float Dataset with the input's calibration, NaN background, FIRE LUT) ----------
struct ThicknessMap {
    std::string name;            // input title + "_Tb.Th" / "_Tb.Sp"
    long w = 0, h = 0, d = 0;
    std::vector<Axis> axes;
    std::vector<float> data;     // [z][y][x]
    double channel_min = 0.0, channel_max = 0.0;   // setChannelMinimum / Maximum (ThicknessWrapper.java:140-141)
    std::string color_table;     // "FIRE"
};

// ---- SharedTable: process-wide results table ------------------------------------------------------------
class SharedTable {
public:
    struct Table {
        std::vector<std::string> headers;                 // column headers
        std::vector<std::string> row_labels;
        std::vector<std::vector<double>> columns;         // [column][row]
        std::vector<std::vector<bool>> filled;            // EMPTY_CELL = not filled
    };
    static void add(const std::string &label, const std::string &header, double value);
    static Table getTable();
    static bool hasData();
    static void reset();

private:
    static std::mutex lock_;
    static Table table_;
};

// ---- ElementUtil / AxisUtils / ResultUtils subsets ------------------------------------------------------
bool isIJ1Binary(const Dataset &ds, long n);
bool has3SpatialDimensions(const Dataset &ds);
bool hasChannelDimensions(const Dataset &ds);
bool hasTimeDimensions(const Dataset &ds);
double getAnisotropy(const Dataset &ds);                   // throws std::invalid_argument like the reference
bool isSpatialCalibrationsIsotropic(const Dataset &ds, double tolerance);
std::string getUnitHeader(const Dataset &ds);              // "(unit)"; ImageJ1 calibrations default to "pixel"

// ---- the command ---------------------------------------------------------------------------------------
class ThicknessWrapper {
public:
    // @Parameter inputs (ThicknessWrapper.java:82-99)
    const Dataset *inputDataset = nullptr;
    std::string mapChoice = "Trabecular thickness";   // | "Trabecular separation" | "Both"
    bool showMaps = true;
    bool maskArtefacts = true;
    // @Parameter outputs (:101-105, BoneJCommand.java:71-72)
    std::unique_ptr<ThicknessMap> trabecularMap;
    std::unique_ptr<ThicknessMap> separationMap;
    SharedTable::Table resultsTable;

    // environment the SciJava context would inject
    bool headless = true;
    bool anisotropyDialogAnswerOk = true;             // what a user answers when not headless
    std::vector<std::string> log;                      // LogService.warn lines, StatusService lines
    int device = 0;

    // CommandService.run(ThicknessWrapper.class, true, ...): validater first, then run()
    void execute();
    bool isCanceled() const { return canceled_; }
    const std::string &getCancelReason() const { return cancelReason_; }

    // the two halves, as in the reference
    void validateImage();                              // ThicknessWrapper.java:225-258
    void run();                                        // :123-156  (throws std::invalid_argument on a bad mapChoice)

private:
    void cancel(const std::string &reason) { canceled_ = true; cancelReason_ = reason; }
    bool canceled_ = false;
    std::string cancelReason_;
    bool anisotropyWarned_ = false;
};

}  // namespace bonej

// ---- C entry points so tests (ctypes) can drive the C++ command ------------------------------------------
extern "C" {
typedef struct bjw_dataset {
    const char *name;
    int n_axes;
    const int *axis_types;        // AxisType as int
    const long *axis_sizes;
    const double *axis_scales;
    const char *const *axis_units;
    int bits_per_pixel;
    const uint8_t *data;
} bjw_dataset;

typedef struct bjw_result {
    int canceled;
    char cancel_reason[256];
    char log[1024];
    float *trabecular_map;        // malloc'ed copies, NULL when the output is null; bjw_result_free releases them
    float *separation_map;
    double trabecular_max, separation_max;
    char color_table[16];
    int n_columns, n_rows;
    char headers[16][64];
    char row_labels[16][128];
    double values[16][16];        // [column][row], NaN where empty
    int exception;                // 1 = IllegalArgumentException (bad mapChoice)
} bjw_result;

int bjw_thickness_run(const bjw_dataset *ds, const char *map_choice, int show_maps, int mask_artefacts, int headless,
                      int dialog_ok, int device, bjw_result *out);
void bjw_result_free(bjw_result *r);
void bjw_shared_table_reset(void);
}
