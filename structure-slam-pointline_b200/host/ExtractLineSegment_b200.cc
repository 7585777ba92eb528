// host/ExtractLineSegment_b200.cc — replaces src/ExtractLineSegment.cpp:18-69 of the reference.  It is compiled
// against the reference's own include/ExtractLineSegment.h (class declaration unchanged), so Frame.cc:152 and
// Tracking.cc:1489 link unchanged.  The reference calls this method through an UNINITIALISED pointer
// (Frame.h:125), so it must not touch `this`: the device handle lives in a function-local thread_local.
#include "ExtractLineSegment.h"
#include "sslpl.h"
#include <stdexcept>
#include <string>
#include <cstring>

namespace StructureSLAM
{
LineSegment::LineSegment() {}

namespace {
struct LineCtx {
    sslpl_line* h; int w, hgt;
    LineCtx(): h(NULL), w(0), hgt(0) {}
    ~LineCtx() { sslpl_line_destroy(h); }
};
}

void LineSegment::ExtractLineSegment(const Mat &img, vector<KeyLine> &keylines, Mat &ldesc, vector<Vector3d> &keylineFunctions, int scale, int numOctaves)
{
    (void)scale; (void)numOctaves;                         // reference passes int(1.2)=1 and 1
    static thread_local LineCtx ctx;
    const int lsdNFeatures = 40;                           // ExtractLineSegment.cpp:42
    if(!ctx.h || img.cols > ctx.w || img.rows > ctx.hgt)
    {
        sslpl_line_destroy(ctx.h); ctx.h = NULL;
        sslpl_line_params p; p.lsdNFeatures = lsdNFeatures;
        p.max_width = img.cols > ctx.w ? img.cols : ctx.w; p.max_height = img.rows > ctx.hgt ? img.rows : ctx.hgt; p.max_batch = 1; p.device = sslpl_default_device();
        if(sslpl_line_create(&p, &ctx.h) != SSLPL_OK)
            throw std::runtime_error(std::string("sslpl_line_create: ") + sslpl_last_error());
        ctx.w = p.max_width; ctx.hgt = p.max_height;
    }
    sslpl_keyline kl[lsdNFeatures]; unsigned char ld[lsdNFeatures*32]; double eq[lsdNFeatures*3];
    int n = 0;
    if(sslpl_line_extract(ctx.h, img.data, img.cols, img.rows, (int)img.step, kl, ld, eq, lsdNFeatures, &n) != SSLPL_OK)
        throw std::runtime_error(std::string("sslpl_line_extract: ") + sslpl_last_error());

    keylines.resize(n);
    for(int i=0; i<n; i++)
    {
        KeyLine &k = keylines[i];
        k.angle = kl[i].angle; k.class_id = kl[i].class_id; k.octave = kl[i].octave;
        k.pt = Point2f(kl[i].pt_x, kl[i].pt_y); k.response = kl[i].response; k.size = kl[i].size;
        k.startPointX = kl[i].startPointX; k.startPointY = kl[i].startPointY; k.endPointX = kl[i].endPointX; k.endPointY = kl[i].endPointY;
        k.sPointInOctaveX = kl[i].sPointInOctaveX; k.sPointInOctaveY = kl[i].sPointInOctaveY;
        k.ePointInOctaveX = kl[i].ePointInOctaveX; k.ePointInOctaveY = kl[i].ePointInOctaveY;
        k.lineLength = kl[i].lineLength; k.numOfPixels = kl[i].numOfPixels;
    }
    ldesc.create(n, 32, CV_8UC1);
    if(n) memcpy(ldesc.data, ld, (size_t)n*32);
    for(int i=0; i<n; i++)                                 // ExtractLineSegment.cpp:56-68 (push_back, no clear)
        keylineFunctions.push_back(Vector3d(eq[3*i], eq[3*i+1], eq[3*i+2]));
}

} // namespace StructureSLAM
