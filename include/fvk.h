/* fvk.h -- C ABI of the B200 (sm_100a) finite-volume kernel library `libfvk.so`.
 *
 * This is the drop-in boundary for the data-parallel hot path of FoamAdapter/NeoN.
 * The reference has no FFI: its "kernels" are C++ lambdas handed to NeoN::parallelFor
 * (src/NeoN/include/NeoN/core/parallelAlgorithms.hpp:29-54) from the bodies of the fvcc operators.
 * Every entry point below replaces the body of one such reference function (cited per function);
 * the C++ host classes in include/NeoN (same names as the reference) call these.
 *
 * Conventions
 *   - plain pointers and sizes only; all array pointers are DEVICE pointers unless the name ends in
 *     `_h` / the comment says host; memory is caller-owned;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); no entry point
 *     synchronises unless documented;
 *   - return value: 0 = success, otherwise an FVK_E* code; fvk_last_error() gives the text;
 *   - there is no CPU fallback: every compute entry point fails with FVK_ENODEVICE without a GPU;
 *   - types follow the reference: scalar = double, label/localIdx = int32_t, Vec3 = 3 contiguous
 *     doubles (AoS, 24 B), sparsity offsets = uint8_t
 *     (src/NeoN/include/NeoN/core/primitives/{scalar,label,vec3}.hpp);
 *   - all fields are in REFERENCE ORDER: cell fields [nCells], face fields [nInternalFaces +
 *     nBoundaryFaces] (internal faces first, then boundary faces in patch order), boundary fields
 *     [nBoundaryFaces].
 */
#ifndef FVK_H
#define FVK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FVK_VERSION 100

enum {
    FVK_OK = 0,
    FVK_EINVAL = 1,    /* bad argument */
    FVK_ENODEVICE = 2, /* no CUDA device / driver */
    FVK_ECUDA = 3,     /* CUDA runtime error, see fvk_last_error */
    FVK_ENOMEM = 4,
    FVK_EUNSUPPORTED = 5,
    FVK_ENCCL = 6
};

typedef void* fvk_stream; /* cudaStream_t */
typedef struct fvk_mesh fvk_mesh;

int fvk_version(void);
/* thread-local text of the last error */
const char* fvk_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * Device memory and copies. Replaces GPUExecutor::alloc/free/realloc (Kokkos::kokkos_malloc,
 * src/NeoN/include/NeoN/core/executor/GPUExecutor.hpp:26-50) and the deep copies behind
 * Vector::copyToHost / copyToExecutor (src/NeoN/src/core/vector/vector.cpp).
 * ---------------------------------------------------------------------------------------------- */
int fvk_device_count(int* n);
int fvk_set_device(int device);
int fvk_malloc(void** dptr, size_t bytes);
int fvk_free(void* dptr);
int fvk_malloc_host(void** hptr, size_t bytes); /* pinned */
int fvk_free_host(void* hptr);
int fvk_memcpy_h2d(void* dst, const void* src_h, size_t bytes, fvk_stream stream);
int fvk_memcpy_d2h(void* dst_h, const void* src, size_t bytes, fvk_stream stream);
int fvk_memcpy_d2d(void* dst, const void* src, size_t bytes, fvk_stream stream);
int fvk_memset(void* dst, int byte, size_t bytes, fvk_stream stream);
int fvk_stream_create(fvk_stream* stream);
int fvk_stream_destroy(fvk_stream stream);
int fvk_stream_sync(fvk_stream stream);

/* ------------------------------------------------------------------------------------------------
 * Mesh description (HOST arrays). This is what readOpenFOAMMesh hands to NeoN::UnstructuredMesh
 * and NeoN::BoundaryMesh (src/datastructures/meshAdapter.cpp:59-136;
 * src/NeoN/include/NeoN/mesh/unstructured/{unstructuredMesh,boundaryMesh}.hpp).
 * nFaces = nInternalFaces + nBoundaryFaces (faces of `empty` patches are dropped, as fvPatch::size()
 * is 0 for them).
 * ---------------------------------------------------------------------------------------------- */
typedef struct fvk_mesh_desc {
    int32_t nCells;
    int32_t nInternalFaces;
    int32_t nBoundaryFaces;
    int32_t nPatches;
    int32_t nPoints;              /* may be 0 */
    const double* points;         /* [nPoints*3] or NULL */
    const double* cellVolumes;    /* [nCells] */
    const double* cellCentres;    /* [nCells*3] */
    const double* faceAreas;      /* Sf [nFaces*3] */
    const double* faceCentres;    /* Cf [nFaces*3] */
    const double* magFaceAreas;   /* [nFaces] */
    const int32_t* faceOwner;     /* [nFaces] (boundary part == faceCells) */
    const int32_t* faceNeighbour; /* [nInternalFaces] */
    /* BoundaryMesh, all [nBoundaryFaces] (x3 for vectors), patch order */
    const int32_t* faceCells;
    const double* bCf;
    const double* bCn;
    const double* bSf;
    const double* bMagSf;
    const double* bNf;
    const double* bDelta;
    const double* bWeights;
    const double* bDeltaCoeffs;
    const int32_t* patchOffsets; /* [nPatches+1] */
    /* decomposed sub-mesh (fvk_decompose): cells [0, nOwnedCells) are owned by this rank, cells
     * [nOwnedCells, nCells) are ghost copies of neighbour ranks' cells; 0 means all cells are owned.
     * Operators produce results for owned cells only. */
    int32_t nOwnedCells;
    /* optional sort key of the internal faces [nInternalFaces] (the global face id of a decomposed
     * mesh): per-cell accumulation follows ascending key instead of ascending local face id, which
     * makes a sub-domain sum in exactly the order of the undecomposed mesh. NULL = local face id. */
    const int32_t* faceOrder;
} fvk_mesh_desc;

/* ------------------------------------------------------------------------------------------------
 * Synthetic single-block hex mesh in OpenFOAM blockMesh ordering with OpenFOAM primitiveMesh
 * geometry -- the stand-in for `blockMesh` + readOpenFOAMMesh, which need OpenFOAM
 * (block (nx ny nz), box lx x ly x lz, simpleGrading 1).
 * Sides: 0 x-min, 1 x-max, 2 y-min, 3 y-max, 4 z-min, 5 z-max. Patches are given in blockMeshDict
 * order, each as an ordered list of sides; `patchIsEmpty[p] != 0` marks an `empty` patch (its faces
 * enter the geometry but not the NeoN mesh). Host only; arrays owned by the returned object.
 * ---------------------------------------------------------------------------------------------- */
int fvk_blockmesh_create(int32_t nx, int32_t ny, int32_t nz, double lx, double ly, double lz,
                         int32_t nPatches, const int32_t* patchNSides, const int32_t* patchSides,
                         const int32_t* patchIsEmpty, int32_t withPoints, fvk_mesh_desc** out);
int fvk_blockmesh_destroy(fvk_mesh_desc* desc);
/* poly faces of the generated mesh (4 point labels per face, all patches incl. empty), for parity
 * checks against polyMesh/faces; valid only if created withPoints. */
int fvk_blockmesh_poly(const fvk_mesh_desc* desc, int32_t* nPolyFaces, const int32_t** facePoints,
                       const int32_t** polyOwner);

/* ------------------------------------------------------------------------------------------------
 * OpenFOAM ASCII polyMesh directories (constant/polyMesh: points, faces, owner, neighbour, boundary) without OpenFOAM:
 * the stand-in for Foam::polyMesh + FoamAdapter::readOpenFOAMMesh (src/datastructures/meshAdapter.cpp:12-45,59-136).
 * Arbitrary polygons / polyhedra; geometry = OpenFOAM primitiveMesh (same arithmetic as the block generator, which
 * reproduces the reference's committed fixtures bit for bit); boundary flattened like the reference's converter (patches
 * in file order, `empty` patches contribute no faces). Host only; arrays owned by the returned object; the description
 * feeds fvk_mesh_create / fvk_decompose like a generated one. Binary-format files are rejected.
 * ---------------------------------------------------------------------------------------------- */
int fvk_polymesh_read(const char* polyMeshDir, fvk_mesh_desc** out);
int fvk_polymesh_destroy(fvk_mesh_desc* desc);
/* name / type of kept (non-empty) patch `patch` of a mesh from fvk_polymesh_read, NUL-terminated into the buffers */
int fvk_polymesh_patch(const fvk_mesh_desc* desc, int32_t patch, char* name, int32_t nameCap, char* type, int32_t typeCap);
/* NOTE: an `empty` patch is dropped from the patch list (the reference keeps it as a zero-size entry that still occupies a
 * patch index, meshAdapter.cpp:33-45,63), so patch indices here count the KEPT patches; this gives kept patch `patch`'s index
 * in the case's boundary file, i.e. the OpenFOAM patch id per-patch lists of a case are keyed by. */
int fvk_polymesh_patch_file_index(const fvk_mesh_desc* desc, int32_t patch, int32_t* fileIndex);
/* write the five files into an existing directory. Faces: faceOffsets [nFaces+1] into facePoints; faces [0, nInternalFaces)
 * are internal, then the patches in order with patchSizes[p] faces each (all patches, `empty` ones included). */
int fvk_polymesh_write(const char* polyMeshDir, int32_t nPoints, const double* points, int32_t nFaces,
                       const int32_t* faceOffsets, const int32_t* facePoints, const int32_t* owner,
                       int32_t nInternalFaces, const int32_t* neighbour, int32_t nPatches,
                       const char* const* patchNames, const char* const* patchTypes, const int32_t* patchSizes);

/* ASCII labelList file with exactly nExpected entries, e.g. constant/cellDecomposition of `decomposePar -cellDist`
 * (scotch, hierarchical, ... -- the reference's cases name several methods in system/decomposeParDict): the cellRank map
 * of fvk_decompose. */
int fvk_labellist_read(const char* path, int32_t nExpected, int32_t* out);
/* OpenFOAM ASCII field files (0/T, 0/U, 0/phi ...) without OpenFOAM, formats as in the reference's fixtures
 * (test/setup_operator/0/): `internalField uniform v | nonuniform List<scalar|vector> n (...)` and the per-patch
 * dictionaries of `boundaryField`. Host only. out (may be NULL to query ncomp) receives nCells * ncomp doubles. */
int fvk_fieldfile_read_internal(const char* path, int32_t nCells, int32_t* ncomp, double* out, int64_t outCapacity);
/* type of patch `patchName` and, when the dictionary has a `value` entry (hasValue = 1), its nPatchFaces * ncomp values
 * (a uniform value is expanded); nPatchFaces < 0 skips the length check of a nonuniform list */
int fvk_fieldfile_read_patch(const char* path, const char* patchName, char* type, int32_t typeCap, int32_t nPatchFaces,
                             int32_t* hasValue, int32_t* ncomp, double* out, int64_t outCapacity);
/* write a vol<Scalar|Vector>Field file (nonuniform internal field, per patch `type` and, where patchHasValue[p] != 0,
 * `value uniform patchValues[p*ncomp..]`); 17 significant digits, so a written field reads back bit for bit */
int fvk_fieldfile_write(const char* path, const char* objectName, int32_t ncomp, int32_t nCells, const double* internal,
                        int32_t nPatches, const char* const* patchNames, const char* const* patchTypes,
                        const int32_t* patchHasValue, const double* patchValues);

/* ------------------------------------------------------------------------------------------------
 * Cell renumbering for coalescing (HOST; BASELINE north_star "renumbered (RCM / space-filling) cell order"). The reference reads
 * whatever order the polyMesh has (meshAdapter.cpp:59-136); OpenFOAM cases are renumbered beforehand with `renumberMesh`. This
 * is that step for mesh descriptions: fvk_renumber_order computes cellOldToNew [nCells] (reverse Cuthill-McKee on the cell-cell
 * graph, or Morton order of the cell centres); fvk_renumber_apply builds the renumbered description (arrays owned by the
 * returned object): cells permuted, internal faces re-sorted into upper-triangular order (owner < neighbour; a face whose
 * owner and neighbour swap gets -Sf), boundary faces and patches unchanged. fvk_renumber_maps gives faceOldToNew [nFaces] and
 * faceFlipped [nFaces] (0/1: a face flux changes sign) to carry fields across. Renumber BEFORE decomposing.
 * ---------------------------------------------------------------------------------------------- */
enum { FVK_RENUMBER_RCM = 0, FVK_RENUMBER_MORTON = 1 };
int fvk_renumber_order(const fvk_mesh_desc* desc_h, int method, int32_t* cellOldToNew);
int fvk_renumber_apply(const fvk_mesh_desc* desc_h, const int32_t* cellOldToNew, fvk_mesh_desc** out);
int fvk_renumber_maps(const fvk_mesh_desc* renumbered, const int32_t** faceOldToNew, const uint8_t** faceFlipped);
int fvk_renumber_destroy(fvk_mesh_desc* renumbered);

/* ------------------------------------------------------------------------------------------------
 * Device mesh handle. Uploads the description and builds, once per mesh:
 *   - BasicGeometryScheme weights / deltaCoeffs / nonOrthDeltaCoeffs
 *     (src/NeoN/src/finiteVolume/cellCentred/stencil/basicGeometryScheme.cpp:15-136),
 *   - the cell->face CSR of CellToFaceStencil::computeStencil
 *     (src/NeoN/src/finiteVolume/cellCentred/stencil/cellToFaceStencil.cpp:14-96),
 *   - the SparsityPattern (src/NeoN/src/linearAlgebra/sparsityPattern.cpp:21-143), bit-exact.
 * ---------------------------------------------------------------------------------------------- */
int fvk_mesh_create(const fvk_mesh_desc* desc_h, fvk_mesh** out);
int fvk_mesh_destroy(fvk_mesh* mesh);

enum fvk_mesh_field {
    /* sizes */
    FVK_N_CELLS = 0, FVK_N_INTERNAL_FACES, FVK_N_BOUNDARY_FACES, FVK_N_PATCHES, FVK_NNZ, FVK_N_OWNED_CELLS,
    /* plan properties (0 / 1): CSR rows laid out like the cell's stencil; block topology proven (index-free kernels) */
    FVK_ROWS_IN_STENCIL_ORDER, FVK_AFFINE_TOPOLOGY,
    /* device arrays, reference order */
    FVK_CELL_VOLUMES = 16, FVK_CELL_CENTRES, FVK_FACE_AREAS, FVK_FACE_CENTRES, FVK_MAG_FACE_AREAS,
    FVK_FACE_OWNER, FVK_FACE_NEIGHBOUR, FVK_FACE_CELLS,
    FVK_B_CF, FVK_B_CN, FVK_B_SF, FVK_B_MAGSF, FVK_B_NF, FVK_B_DELTA, FVK_B_WEIGHTS,
    FVK_B_DELTACOEFFS,
    FVK_WEIGHTS = 48, FVK_DELTACOEFFS, FVK_NONORTH_DELTACOEFFS,  /* [nFaces] */
    FVK_STENCIL_SEGMENTS = 64, /* int32 [nCells+1] */
    FVK_STENCIL_VALUES,        /* int32 [2*nI+nB], sorted face ids per cell */
    FVK_ROW_OFFS = 80,         /* int32 [nCells+1] */
    FVK_COL_IDXS,              /* int32 [nnz] */
    FVK_OWNER_OFFSET,          /* uint8 [nI]  -> entry (row own, col nei) */
    FVK_NEIGHBOUR_OFFSET,      /* uint8 [nI]  -> entry (row nei, col own) */
    FVK_DIAG_OFFSET            /* uint8 [nCells] */
};
/* integer properties (FVK_N_*, FVK_NNZ) */
int fvk_mesh_size(const fvk_mesh* mesh, int field, int64_t* value);
/* device pointer + element count of an array field */
int fvk_mesh_array(const fvk_mesh* mesh, int field, const void** dptr, int64_t* count);
/* host copy of patchOffsets [nPatches+1] */
int fvk_mesh_patch_offsets(const fvk_mesh* mesh, int32_t* offsets_h);

/* ------------------------------------------------------------------------------------------------
 * Explicit operators. One fused, deterministic, cell-centric gather kernel
 * per operator: no atomics, no face-sized temporary, per-cell summation in ascending face id = the
 * order of the reference's SerialExecutor.
 *
 * coeff, coeffView : dsl::Coeff (src/NeoN/include/NeoN/dsl/coeff.hpp:35); operatorScaling[c] =
 *                    coeffView ? coeffView[c]*coeff : coeff.
 * mode             : FVK_SET        out[c]  = sum * s[c]           (out assumed zero on entry)
 *                    FVK_ACC_SCALE  out[c]  = (out[c] + sum) * s[c] (computeDiv on a non-zero `res`)
 *                    FVK_ADD        out[c] += sum * s[c]            (Operator::explicitOperation:
 *                                   tmp=0; op(tmp); source += tmp,  operators/divOperator.hpp:133-140)
 *                    with s[c] = operatorScaling[c] / V[c].
 * ---------------------------------------------------------------------------------------------- */
enum { FVK_SET = 0, FVK_ACC_SCALE = 1, FVK_ADD = 2 };
enum { FVK_LINEAR = 0, FVK_UPWIND = 1 };

/* GaussGreenDiv<scalar|Vec3>::div -> computeDivExp + computeDiv
 * (src/NeoN/src/finiteVolume/cellCentred/operators/gaussGreenDiv.cpp:29-140) fused with the face
 * interpolation (interpolation/linear.cpp:12-46, upwind.cpp:12-56). */
int fvk_div_s(const fvk_mesh* mesh, int scheme, const double* faceFlux, const double* phi,
              const double* phiBValue, double coeff, const double* coeffView, double* out, int mode,
              fvk_stream stream);
int fvk_div_v(const fvk_mesh* mesh, int scheme, const double* faceFlux, const double* phi,
              const double* phiBValue, double coeff, const double* coeffView, double* out, int mode,
              fvk_stream stream);
/* GaussGreenGrad::grad -> computeGrad (operators/gaussGreenGrad.cpp:18-71); always linear, scale 1/V.
 * mode FVK_SET or FVK_ACC_SCALE. out is Vec3[nCells]. */
int fvk_grad_s(const fvk_mesh* mesh, const double* phi, const double* phiBValue, double* out,
               int mode, fvk_stream stream);
/* GaussGreenLaplacian::laplacian -> computeLaplacianExp (operators/gaussGreenLaplacian.cpp:11-61)
 * fused with computeFaceNormalGrad (faceNormalGradient/uncorrected.cpp:11-52). gamma is ignored by
 * the reference (parameter unnamed, gaussGreenLaplacian.cpp:14) and therefore not an argument. */
int fvk_laplacian_s(const fvk_mesh* mesh, const double* phi, const double* phiBValue, double coeff,
                    const double* coeffView, double* out, int mode, fvk_stream stream);
int fvk_laplacian_v(const fvk_mesh* mesh, const double* phi, const double* phiBValue, double coeff,
                    const double* coeffView, double* out, int mode, fvk_stream stream);
/* surfaceIntegrate (operators/surfaceIntegrate.cpp:11-49): scatter of a given face flux. */
int fvk_surface_integrate_s(const fvk_mesh* mesh, const double* flux, double coeff,
                            const double* coeffView, double* out, int mode, fvk_stream stream);
int fvk_surface_integrate_v(const fvk_mesh* mesh, const double* flux, double coeff,
                            const double* coeffView, double* out, int mode, fvk_stream stream);

/* SurfaceInterpolation::interpolate (interpolation/linear.cpp:12-46, upwind.cpp:12-56).
 * faceFlux may be NULL for FVK_LINEAR. outFace is [nFaces] (x3 for _v). */
int fvk_interpolate_s(const fvk_mesh* mesh, int scheme, const double* faceFlux, const double* phi,
                      const double* phiBValue, double* outFace, fvk_stream stream);
int fvk_interpolate_v(const fvk_mesh* mesh, int scheme, const double* faceFlux, const double* phi,
                      const double* phiBValue, double* outFace, fvk_stream stream);
/* SurfaceInterpolation::weight (linear.hpp:76-96 copy of the geometric weights; upwind.cpp:59-93).
 * wFace [nFaces], wBoundary [nBoundaryFaces]. */
int fvk_interpolation_weights(const fvk_mesh* mesh, int scheme, const double* faceFlux,
                              double* wFace, double* wBoundary, fvk_stream stream);
/* FaceNormalGradient Uncorrected::faceNormalGrad (faceNormalGradient/uncorrected.cpp:11-52). */
int fvk_face_normal_grad_s(const fvk_mesh* mesh, const double* phi, const double* phiBValue,
                           double* outFace, fvk_stream stream);
int fvk_face_normal_grad_v(const fvk_mesh* mesh, const double* phi, const double* phiBValue,
                           double* outFace, fvk_stream stream);

/* computeCoNum (src/NeoN/src/finiteVolume/cellCentred/auxiliary/coNum.cpp:18-96).
 * result_d[0] = maxCoNum, result_d[1] = meanCoNum (device, 2 doubles). scratch_d: >= fvk_conum_scratch
 * bytes. Deterministic two-stage reduction; no host sync. */
size_t fvk_conum_scratch_bytes(const fvk_mesh* mesh);
int fvk_conum(const fvk_mesh* mesh, const double* faceFlux, double dt, double* result_d,
              void* scratch_d, fvk_stream stream);

/* Volume boundary conditions, one launch for all patches of a field
 * (boundary/volume/fixedValue.hpp:21-43, fixedGradient.hpp:22-54, extrapolated.hpp:22-53,
 * calculated/empty: no-op). Per patch p: kind[p], and value[p*ncomp..] = the fixedValue /
 * fixedGradient constant. Arrays kind_h/value_h are HOST (tiny, copied as kernel arguments;
 * nPatches <= FVK_MAX_PATCHES). ncomp = 1 (scalar) or 3 (Vec3). */
enum { FVK_BC_CALCULATED = 0, FVK_BC_FIXED_VALUE = 1, FVK_BC_FIXED_GRADIENT = 2,
       FVK_BC_EXTRAPOLATED = 3, FVK_BC_EMPTY = 4 };
#define FVK_MAX_PATCHES 16
int fvk_correct_boundary_conditions(const fvk_mesh* mesh, int ncomp, const int32_t* kind_h,
                                    const double* value_h, const double* internal, double* bValue,
                                    double* bRefValue, double* bValueFraction, double* bRefGrad,
                                    fvk_stream stream);

/* ------------------------------------------------------------------------------------------------
 * Implicit assembly into LinearSystem<scalar|Vec3, localIdx> (values T[nnz], rhs T[nCells]) over the
 * mesh's SparsityPattern, plus the BoundaryCoefficients arrays (linearSystem.hpp:36-43).
 *
 * One fused kernel assembles an ordered list of terms; row c is written by one thread, each entry
 * exactly once, with the accumulation order of applying the reference operators one after another
 * (Expression::implicitOperation: spatial operators in insertion order, then temporal,
 * src/NeoN/include/NeoN/dsl/expression.hpp:80-101; timeIntegration/backwardEuler.hpp:49-50).
 *   FVK_TERM_DDT        DdtOperator::implicitOperation   (operators/ddtOperator.cpp:38-60)
 *                       cellField = old-time field (T[nCells]), dt
 *   FVK_TERM_DIV        computeDivImp                    (operators/gaussGreenDiv.cpp:155-262)
 *                       faceField = faceFlux [nFaces], scheme = FVK_LINEAR | FVK_UPWIND
 *   FVK_TERM_LAPLACIAN  computeLaplacianImpl             (operators/gaussGreenLaplacian.cpp:76-177)
 *                       faceField = gamma [nFaces] (or gammaCell / gammaBoundary, see fvk_term)
 *   FVK_TERM_SOURCE     SourceTerm::implicitOperation    (operators/sourceTerm.cpp:37-55)
 *                       cellField = coefficients (double[nCells])
 * coeff/coeffView = the operator's dsl::Coeff (a subtracted operator carries coeff = -1,
 * dsl/expression.hpp:207-224).
 * accumulate = 0: write a fresh system (replaces createEmptyLinearSystem's zero-fill + the operators);
 * accumulate = 1: add to the existing values/rhs (an operator applied to a non-empty system).
 * bcMatrix/bcRhs [nBoundaryFaces] receive the boundary coefficients of the LAST div/laplacian term
 * (each reference operator overwrites them, gaussGreenDiv.cpp:251,258).
 * ---------------------------------------------------------------------------------------------- */
enum { FVK_TERM_DDT = 0, FVK_TERM_DIV = 1, FVK_TERM_LAPLACIAN = 2, FVK_TERM_SOURCE = 3 };
#define FVK_MAX_TERMS 6
typedef struct fvk_term {
    int32_t kind;
    int32_t scheme;
    double coeff;
    const double* coeffView; /* [nCells] or NULL */
    const double* faceField; /* div: faceFlux, laplacian: gamma; [nFaces] */
    const double* cellField; /* ddt: old field T[nCells]; source: coefficients double[nCells] */
    double dt;
    /* laplacian only, faceField == NULL: gamma is the LINEAR INTERPOLATE of this cell field [nCells] / its boundary values
     * [nBoundaryFaces], evaluated on the fly with the arithmetic of computeLinearInterpolation (interpolation/linear.cpp:30-45)
     * -- neoIcoFoam's laplacian(interpolate(rAU), p) (neoIcoFoam.cpp:117-124,141) without the face-sized rAUf temporary */
    const double* gammaCell;
    const double* gammaBoundary;
} fvk_term;
/* BoundaryData of the unknown field (fields/boundaryData.hpp:32-215), device pointers */
typedef struct fvk_bfield {
    const double* value;         /* T[nB] */
    const double* refValue;      /* T[nB] */
    const double* valueFraction; /* double[nB] */
    const double* refGrad;       /* T[nB] */
} fvk_bfield;
int fvk_assemble_s(const fvk_mesh* mesh, int nTerms, const fvk_term* terms_h, const fvk_bfield* bd,
                   double* values, double* rhs, double* bcMatrix, double* bcRhs, int accumulate,
                   fvk_stream stream);
int fvk_assemble_v(const fvk_mesh* mesh, int nTerms, const fvk_term* terms_h, const fvk_bfield* bd,
                   double* values, double* rhs, double* bcMatrix, double* bcRhs, int accumulate,
                   fvk_stream stream);
/* Compact Vec3 system (new here; an HBM-layout choice, not a reference type): every implicit operator multiplies its
 * coefficient by one<Vec3>() (gaussGreenDiv.cpp:200,209), so the three components of a matrix entry are
 * identical -- valuesCompact double[nnz] stores each entry once (a third of the traffic of LinearSystem<Vec3>::values);
 * rhs, bcMatrix and bcRhs stay Vec3. fvk_expand_vec3 materialises the reference layout Vec3[nnz] from it, bit for bit. */
int fvk_assemble_vc(const fvk_mesh* mesh, int nTerms, const fvk_term* terms_h, const fvk_bfield* bd,
                    double* valuesCompact, double* rhs, double* bcMatrix, double* bcRhs, int accumulate,
                    fvk_stream stream);
int fvk_expand_vec3(int64_t n, const double* compact, double* outV, fvk_stream stream);
/* BoundaryCoefficients::matrixIdxs / rhsIdxs as createEmptyLinearSystem fills them
 * (linearSystem.hpp:163-174; matrixIdxs = celli + diagOffset[celli], sic). */
int fvk_bc_coeff_indices(const fvk_mesh* mesh, int32_t* matrixIdxs, int32_t* rhsIdxs,
                         fvk_stream stream);
/* DdtOperator::explicitOperation (ddtOperator.cpp:21-36): source += (field - old)/dt * V */
int fvk_ddt_explicit(const fvk_mesh* mesh, int ncomp, const double* field, const double* oldField,
                     double dt, double* source, fvk_stream stream);
/* SourceTerm::explicitOperation (sourceTerm.cpp:22-35): source += coeff * k * field */
int fvk_source_explicit(const fvk_mesh* mesh, int ncomp, const double* k, const double* field,
                        double coeff, const double* coeffView, double* source, fvk_stream stream);
/* dsl::solve (dsl/solver.hpp:73-77): rhs -= explicitSource * V */
int fvk_rhs_sub_source(const fvk_mesh* mesh, int ncomp, const double* src, double* rhs,
                       fvk_stream stream);

/* ------------------------------------------------------------------------------------------------
 * Linear algebra.
 * ---------------------------------------------------------------------------------------------- */
/* CSR y = A x over LinearSystem<scalar,localIdx> arrays (CSRMatrix.hpp:19-74). Rows are summed in
 * ascending entry order like the reference's row loop (linearAlgebra/utilities.cpp:21-34). x has
 * nCols >= nRows entries (ghost columns of a decomposed mesh follow the owned rows). */
int fvk_spmv(int32_t nRows, const int32_t* rowOffs, const int32_t* colIdxs, const double* values,
             const double* x, double* y, fvk_stream stream);
/* computeResidual (src/NeoN/src/linearAlgebra/utilities.cpp:11-35): res = A x - b */
/* y = A x over the mesh's own SparsityPattern (la::SparsityPattern(mesh)), structured fast path as in
 * fvk_solver_attach_mesh; otherwise identical to fvk_spmv(mesh rows, mesh rowOffs, mesh colIdxs, ...). */
int fvk_spmv_structured(const fvk_mesh* mesh, const double* values, const double* x, double* y, fvk_stream stream);
int fvk_residual(int32_t nRows, const int32_t* rowOffs, const int32_t* colIdxs, const double* values,
                 const double* b, const double* x, double* res, fvk_stream stream);
/* Vector free functions (src/NeoN/src/core/vector/vectorFreeFunctions.cpp:18-106,
 * core/containerFreeFunctions.hpp:51-115): fill, setContainer (copy = fvk_memcpy_d2d), scalarMul,
 * add / sub / mul (x op= y), and the axpby the time integrators spell as temporaries
 * (forwardEuler.hpp:44-49: phi = old - source*dt). n counts doubles (3*n for a Vec3 vector). */
int fvk_vec_fill(int64_t n, double value, double* x, fvk_stream stream);
int fvk_vec_scale(int64_t n, double a, double* x, fvk_stream stream);
int fvk_vec_add(int64_t n, double* x, const double* y, fvk_stream stream);
int fvk_vec_sub(int64_t n, double* x, const double* y, fvk_stream stream);
int fvk_vec_mul(int64_t n, double* x, const double* y, fvk_stream stream);
/* y = a*x + b*y */
int fvk_vec_axpby(int64_t n, double a, const double* x, double b, double* y, fvk_stream stream);
/* w = a*x + b*y (w may alias x or y): forwardEuler.hpp:48 `solution = old - source * dt` in one pass */
int fvk_vec_waxpby(int64_t n, double a, const double* x, double b, const double* y, double* w, fvk_stream stream);
/* out = x * a: the temporary `Vector operator*(Vector, scalar)` creates (vector.hpp; scalarAdvection.cpp:66-67
 * nfPhi = nfPhi0 * cos(...)), written straight to its destination */
int fvk_vec_scaled_copy(int64_t n, double a, const double* x, double* out, fvk_stream stream);
/* result_d[0] = sum x[i]*y[i]  /  sqrt(sum x[i]^2): deterministic two-stage warp-shuffle reduction,
 * result stays on the device (no host sync). */
int fvk_dot(int64_t n, const double* x, const double* y, double* result_d, fvk_stream stream);
int fvk_norm2(int64_t n, const double* x, double* result_d, fvk_stream stream);

/* la::Solver (linearAlgebra/solver.hpp:14-91) with the Ginkgo configuration mapFvSolution produces
 * for the PISO pressure equation (src/compatibility/fvSolution.cpp:19-159): solver::Cg, optional
 * preconditioner::Jacobi (max_block_size 1), criteria iteration / relative_residual_norm /
 * absolute_residual_norm OR-combined. Ginkgo itself is a third-party dependency of the reference
 * (1.10.0, src/NeoN/cmake/CxxThirdParty.cmake:158-179); the iteration follows its published CG
 *  and GinkgoSolver::solve's bookkeeping (linearAlgebra/ginkgo.hpp:116-155):
 * initResNorm = ||b||_2 (sic), finalResNorm = ||r||_2 at stop, numIter = completed updates.
 * Two kernels per iteration: {x += alpha p; r -= alpha q; z = r/d; r.z; r.r} and
 * {p = z + beta p; q = A p; p.q}; scalars stay on the device; the host polls the stop flag every
 * `checkEvery` iterations (kernels after the stop are no-ops, so results equal a per-iteration check). */
typedef struct fvk_solver fvk_solver;
typedef struct fvk_comm fvk_comm;
/* FVK_PRECOND_DIC (extension; the reference maps OpenFOAM's DIC to scalar Jacobi, fvSolution.cpp:51-55): diagonal
 * incomplete Cholesky, OpenFOAM's DIC recurrences applied in a MULTICOLOUR order of the cells (greedy colouring of the mesh's own
 * pattern; a hex block needs two colours), so every substitution step is one fully parallel kernel. solver::Cg on one GPU, needs
 * fvk_solver_attach_mesh. Changes iteration counts by construction: compared with the oracle's restatement of the same ordering. */
enum { FVK_PRECOND_NONE = 0, FVK_PRECOND_JACOBI = 1, FVK_PRECOND_DIC = 2 };
/* solver::Cg (PCG) | solver::Bicgstab (PBiCGStab, smoothSolver: src/compatibility/fvSolution.cpp:22-28). BiCGStab follows
 * Ginkgo 1.10's published loop (two stopping checks per iteration, on ||r|| and on the half-step residual ||s||; a stop at
 * the half step adds alpha*y to x and does not count as an iteration); five kernels per iteration, scalars on the device. */
enum { FVK_SOLVER_CG = 0, FVK_SOLVER_BICGSTAB = 1 };
typedef struct fvk_solver_config {
    int32_t maxIter;        /* criteria.iteration (default 1000, fvSolution.cpp:117-138) */
    double relTol;          /* criteria.relative_residual_norm */
    double absTol;          /* criteria.absolute_residual_norm */
    int32_t preconditioner; /* FVK_PRECOND_* */
    int32_t checkEvery;     /* host polls the device stop flag every this many iterations (>=1) */
    int32_t solverType;     /* FVK_SOLVER_* */
} fvk_solver_config;
typedef struct fvk_solver_stats {
    int32_t numIter;
    double initResNorm;
    double finalResNorm;
    int32_t nHistory; /* entries written to history_h */
} fvk_solver_stats;
/* nRows owned rows, nCols >= nRows entries in x (ghosts after the owned rows). comm may be NULL
 * (single GPU); with a comm, dots are all-reduced and p is halo-exchanged before every A p. */
int fvk_solver_create(int32_t nRows, int32_t nCols, const fvk_solver_config* cfg, fvk_comm* comm,
                      fvk_solver** out);
int fvk_solver_destroy(fvk_solver* solver);
/* Solves A x = b in place (x = initial guess). history_h (HOST, may be NULL) receives ||r||_2 at
 * every stopping check (history[0] = ||r0||), at most maxHistory entries. Synchronises `stream`. */
/* Optional: tell the solver which mesh its systems come from (the mesh's own SparsityPattern arrays will be passed to
 * fvk_solver_solve). When the mesh plan proved a block-structured topology, the SpMV inside CG computes the column indices
 * of the regular rows instead of reading them (28 of 104 bytes per row); results are bit-identical. NULL detaches. */
int fvk_solver_attach_mesh(fvk_solver* solver, const fvk_mesh* mesh);
int fvk_solver_solve(fvk_solver* solver, const int32_t* rowOffs, const int32_t* colIdxs,
                     const double* values, const double* b, double* x, fvk_solver_stats* stats_h,
                     double* history_h, int32_t maxHistory, fvk_stream stream);

/* fvk_solver_solve on a CAPTURING stream (cudaStreamBeginCapture / torch.cuda.graph): solver::Cg on one GPU or over the
 * peer-memory transport becomes part of the graph being captured -- the start-up kernels, then a conditional WHILE node whose
 * body is two CG iterations and whose condition a device kernel clears when the stopping criterion fires. No host round trip:
 * a whole PISO step (assembly, both pressure solves, correctors) is ONE graph launch. stats.numIter = -(slot + 1) at capture
 * time; the real statistics of a replay are read with fvk_solver_captured_stats(slot) once the replay has completed. */
int fvk_solver_captured_stats(const fvk_solver* solver, int32_t slot, fvk_solver_stats* stats_h);
int fvk_solver_reset_captures(fvk_solver* solver);
/* iteration counts of the captured solves executed since the last call, in execution order (a device-side log, so that graph
 * replays issued back to back need no host synchronisation to keep their statistics); synchronises the device */
int fvk_solver_captured_log(fvk_solver* solver, int32_t* out_h, int32_t capacity, int32_t* n_h);
/* Ghost entries around a distributed solve (a solver created with a communicator; no-ops without one).
 * fvk_solver_set_ghosts_current(on): the caller guarantees that the ghost entries of every initial guess passed from now on are
 * current (e.g. the previous solution of the same field, untouched since), so the solver skips its start-up exchange of x.
 * fvk_solver_keeps_ghosts: *out_h = 1 when the solver leaves the ghost entries of its SOLUTION current (solver::Cg over the
 * peer-memory transport: the ghost entries follow x += alpha p with the exchanged search direction, bit-identical to the
 * owners' values), so no exchange of x is needed after the solve; 0 otherwise (NCCL transport, BiCGStab). */
int fvk_solver_set_ghosts_current(fvk_solver* solver, int32_t on);
int fvk_solver_keeps_ghosts(const fvk_solver* solver, int32_t* out_h);
/* Vec3 LinearSystem (values Vec3[nnz] with identical components; rhs / x Vec3[nRows] / [nCols]): three scalar
 * solves over the component matrix (NeoN's la::Solver has no Vec3 overload, solver.hpp:52; this is what `momentumPredictor yes`
 * of neoIcoFoam.cpp:100-103 needs). stats3_h receives one fvk_solver_stats per component. */
int fvk_solver_solve_vec3(fvk_solver* solver, int64_t nnz, const int32_t* rowOffs, const int32_t* colIdxs, const double* valuesV,
                          const double* bV, double* xV, fvk_solver_stats* stats3_h, fvk_stream stream);
/* same with the compact matrix of fvk_assemble_vc (double[nnz], used in place: no component extraction) */
int fvk_solver_solve_vec3c(fvk_solver* solver, const int32_t* rowOffs, const int32_t* colIdxs, const double* valuesCompact,
                           const double* bV, double* xV, fvk_solver_stats* stats3_h, fvk_stream stream);

/* ------------------------------------------------------------------------------------------------
 * PISO pressure-velocity coupling (FoamAdapter src/algorithms/pressureVelocityCoupling.cpp) and the
 * PDESolver helpers (include/FoamAdapter/datastructures/expression.hpp). valuesV / rhsV are the Vec3
 * momentum LinearSystem (identical components; component [0] is read, like the reference).
 * ---------------------------------------------------------------------------------------------- */
/* computeRAU (:38-63) and computeRAUandHByA (:65-128) in one cell-centric pass:
 * rAU[c] = V[c] / diag[c][0];  HbyA[c] = (rhs[c] - sum_offdiag A[c][k][0] * U[k]) * (rAU[c] / V[c]),
 * off-diagonals subtracted in ascending face order. HbyA == NULL computes rAU only.
 * Boundary values follow from fvk_correct_boundary_conditions with FVK_BC_EXTRAPOLATED. */
int fvk_rAU_HbyA(const fvk_mesh* mesh, const double* valuesV, const double* rhsV, const double* U,
                 double* rAU, double* HbyA, fvk_stream stream);
/* same, from the compact momentum matrix of fvk_assemble_vc (needs a mesh whose CSR rows are in stencil order) */
int fvk_rAU_HbyA_c(const fvk_mesh* mesh, const double* valuesCompact, const double* rhsV, const double* U,
                   double* rAU, double* HbyA, fvk_stream stream);
/* constrainHbyA (:14-36): dstB = srcB on the patches with patchMask_h[p] != 0 (the non-assignable
 * patches of U). patchMask_h is HOST [nPatches]. */
int fvk_copy_patches(const fvk_mesh* mesh, int ncomp, const int32_t* patchMask_h, const double* srcB,
                     double* dstB, fvk_stream stream);
/* flux (:215-267): outFace[f] = Sf . (w (U_P - U_N) + U_N); boundary bSf . Ub, also to outB (may be NULL) */
int fvk_flux(const fvk_mesh* mesh, const double* U, const double* Ub, double* outFace, double* outB,
             fvk_stream stream);
/* updateFaceVelocity (:131-197) from the assembled scalar pressure system + its boundary coefficients */
int fvk_update_face_velocity(const fvk_mesh* mesh, const double* values, const double* bcMatrix,
                             const double* bcRhs, const double* p, const double* predPhi,
                             const double* predPhiB, double* phi, double* phiB, fvk_stream stream);
/* updateVelocity (:199-213): U = HbyA - rAU * gradP (gradP from fvk_grad_s) */
int fvk_update_velocity(const fvk_mesh* mesh, const double* HbyA, const double* rAU, const double* gradP,
                        double* U, fvk_stream stream);
/* updateVelocity fused with the gradient it consumes: U = HbyA - rAU * grad(p) in one pass (no gradP array in memory); same
 * arithmetic, bit for bit, as fvk_grad_s (FVK_SET) followed by fvk_update_velocity */
int fvk_update_velocity_grad(const fvk_mesh* mesh, const double* HbyA, const double* rAU, const double* p, const double* pB,
                             double* U, fvk_stream stream);
/* forwardEuler (timeIntegration/forwardEuler.hpp:38-56: solution = old - source * dt) of `ddt(phi) + div(faceFlux, phi)` with the div
 * as its only spatial operator (examples/scalarAdvection/scalarAdvection.cpp:77-83) in ONE pass: out = phiOld - dt * (coeff / V) *
 * sum_f flux_f phi_f, bit for bit fvk_div_s (FVK_SET) into a source followed by fvk_vec_waxpby(-dt, source, 1, phiOld, out), without
 * the source vector. phiOld (owned + ghost entries) and out must be different arrays. */
int fvk_div_forward_euler_s(const fvk_mesh* mesh, int scheme, const double* faceFlux, const double* phiOld, const double* phiB,
                            double coeff, const double* coeffView, double dt, double* out, fvk_stream stream);
/* dsl::solve's explicit source when it is ONE surfaceIntegrate operator (the pressure equation's `- exp::div(phiHbyA)`):
 * rhs -= (0 + coeff/V * sum_f flux_f) * V in one pass = SurfaceIntegrate::explicitOperation into a zeroed source
 * (operators/surfaceIntegrate.hpp) followed by dsl/solver.hpp:73-77, bit for bit, without the source vector */
int fvk_rhs_sub_surface_integrate_s(const fvk_mesh* mesh, const double* flux, double coeff, const double* coeffView,
                                    double* rhs, fvk_stream stream);
/* PDESolver::SetReference (expression.hpp:86-112): rhs[r] += diag*value; diag += diag */
int fvk_set_reference(const fvk_mesh* mesh, int32_t refCell, double refValue, double* values, double* rhs,
                      fvk_stream stream);
/* diag(ls, sparsityPattern) (expression.hpp:181-199) */
int fvk_diag(const fvk_mesh* mesh, int ncomp, const double* values, double* out, fvk_stream stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU: one process per GPU, NCCL over NVLink/NVSwitch. Replaces the reference's unhooked MPI
 * halo layer (src/NeoN/include/NeoN/mesh/unstructured/communicator.hpp:89-143 startComm /
 * finaliseComm; core/mpi/operators.hpp:132-160 allReduce), which packs and unpacks on the host.
 * A decomposed sub-mesh keeps its ghost cells directly behind its nOwned cells in every cell field
 * (see fvk_decompose), so received values land in place and every kernel indexes them like cells.
 * ---------------------------------------------------------------------------------------------- */
#define FVK_UNIQUE_ID_BYTES 128
/* rank 0 creates the id; the host distributes it (torch.distributed / MPI broadcast) */
int fvk_comm_unique_id(void* id128);
/* collective over all ranks (ncclCommInitRank on the current device); nRanks == 1 needs no id */
int fvk_comm_create(int rank, int nRanks, const void* id128, fvk_comm** out);
int fvk_comm_destroy(fvk_comm* comm);
int fvk_comm_rank(const fvk_comm* comm, int* rank, int* nRanks);
/* halo plan (HOST arrays): for neighbour k, sendCells[sendOffsets[k]..sendOffsets[k+1]) are owned
 * cells whose values rank neighbourRanks[k] needs, in the order of that rank's ghost slots; this
 * rank's ghosts owned by neighbour k are field[nOwned + recvOffsets[k] .. nOwned + recvOffsets[k+1]). */
int fvk_comm_set_halo(fvk_comm* comm, int32_t nOwned, int32_t nNeighbours, const int32_t* neighbourRanks_h,
                      const int32_t* sendOffsets_h, const int32_t* sendCells_h, const int32_t* recvOffsets_h);
/* pack + one NCCL send/recv group on `stream`; field has (nOwned + nGhost) * ncomp doubles */
int fvk_comm_halo_exchange(fvk_comm* comm, double* field, int ncomp, fvk_stream stream);
/* several cell fields (up to 4, at most 8 components together) in ONE exchange: one push kernel, one flag per neighbour, one
 * wait (peer-memory transport; over NCCL it is one send/recv group per field). PISO exchanges rAU + HbyA this way. */
typedef struct fvk_halo_field { double* field; int32_t ncomp; } fvk_halo_field;
int fvk_comm_halo_exchange_multi(fvk_comm* comm, int nFields, const fvk_halo_field* fields_h, fvk_stream stream);
int fvk_comm_allreduce_sum(fvk_comm* comm, double* data_d, int count, fvk_stream stream);
int fvk_comm_allreduce_max(fvk_comm* comm, double* data_d, int count, fvk_stream stream);
/* Peer-memory transport (NVLink/NVSwitch P2P through CUDA IPC), optional, after fvk_comm_set_halo: every rank exports
 * a window blob (FVK_P2P_BLOB_BYTES), the host all-gathers the blobs (torch.distributed / MPI_Allgather), every rank
 * connects. From then on fvk_comm_halo_exchange stores the send cells straight into the neighbours' windows and raises
 * a flag (no NCCL kernel), fvk_comm_allreduce_sum (count <= 4) is a mailbox exchange, and the distributed CG fuses its
 * all-reduces into the kernels that produce the partial sums. Sums are formed in rank order on every rank
 * (bit-identical everywhere, run to run). fvk_comm_p2p_connect fails (NCCL stays the transport) when CUDA IPC is
 * unavailable. All ranks must issue the same sequence of exchanges / reductions. */
#define FVK_P2P_BLOB_BYTES 512
int fvk_comm_p2p_export(fvk_comm* comm, void* blob_h);
int fvk_comm_p2p_connect(fvk_comm* comm, const void* allBlobs_h /* nRanks x FVK_P2P_BLOB_BYTES, rank order */);
int fvk_comm_p2p_enabled(const fvk_comm* comm);
/* back to NCCL; the transport must be the same on every rank, so call it everywhere when any rank failed to connect */
int fvk_comm_p2p_disable(fvk_comm* comm);
/* diagnostics: accumulated ns spent by the CG kernels' last blocks in [0] flag raise, [1] all-reduce (r.z, r.r), [2] halo
 * flag wait, [3] count, [4] all-reduce p.q, [5] count; [6] ns the field exchanges waited for their neighbours' flags, [7] count */
int fvk_comm_p2p_debug(const fvk_comm* comm, uint64_t* out8_h);

/* ------------------------------------------------------------------------------------------------
 * Domain decomposition (HOST). The reference carries OpenFOAM decomposeParDict files
 * (tutorials/cavity/system/decomposeParDict:17-24, method hierarchical/simple, n (px py pz)) but no
 * decomposed solver path, so the layout is defined here: a sub-mesh keeps GHOST CELLS
 * behind its owned cells (grouped by owning rank, ascending global id), every global internal face that
 * touches an owned cell stays an internal face (faceOrder = global face id, so sub-domain sums follow the
 * undecomposed mesh's order), and boundary faces keep their patch. See fvk_decomp.cpp.
 * ---------------------------------------------------------------------------------------------- */
typedef struct fvk_decomp fvk_decomp;
/* OpenFOAM `simple` restated: per axis, stable sort by cell-centre coordinate, p equal-count groups;
 * rank = bx + px*(by + py*bz). cellRank [nCells] (host, out). */
int fvk_decomp_simple_map(const fvk_mesh_desc* global_h, int px, int py, int pz, int32_t* cellRank);
int fvk_decompose(const fvk_mesh_desc* global_h, const int32_t* cellRank, int nRanks, int rank,
                  fvk_decomp** out);
int fvk_decomp_destroy(fvk_decomp* d);
/* the sub-mesh description (owned by the fvk_decomp); feed it to fvk_mesh_create */
const fvk_mesh_desc* fvk_decomp_mesh(const fvk_decomp* d);
int fvk_decomp_info(const fvk_decomp* d, int32_t* nOwned, int32_t* nGhost, int32_t* nNeighbours);
/* local -> global maps: cellGlobal [nOwned+nGhost]; faceGlobal [local nI + nB] (global face id, boundary
 * faces as nI_global + global boundary index) */
int fvk_decomp_maps(const fvk_decomp* d, const int32_t** cellGlobal, const int32_t** faceGlobal);
/* halo plan in the form fvk_comm_set_halo takes */
int fvk_decomp_halo(const fvk_decomp* d, const int32_t** neighbourRanks, const int32_t** sendOffsets,
                    const int32_t** sendCells, const int32_t** recvOffsets);
int fvk_comm_set_halo_from_decomp(fvk_comm* comm, const fvk_decomp* d);

/* experiment switch: selects the kernel variant used by the gather operators. 0 = default (the brick
 * kernel when the mesh has a brick plan, else the per-cell gather), 5 = per-cell gather, 1-4 = packed-plan
 * gathers, 6 = TMA tile kernel, 7 = brick kernel. Used by the roofline harness and the parity tests only. */
int fvk_set_variant(int variant);
/* Halo overlap for the cell-centric explicit operators (div / grad / laplacian / surface_integrate) on a decomposed
 * mesh: phase FVK_TILES_INTERIOR makes them compute only the cells of tiles that read no ghost cell (safe to run
 * while the halo exchange of the operand is in flight), FVK_TILES_HALO only the remaining tiles (after the exchange);
 * FVK_TILES_ALL (default) computes everything. The two phases together write every owned cell exactly once, with
 * bit-identical results. New here: the reference has no distributed operator path (communicator.hpp:89-143 is the
 * host-side start/finish protocol this mirrors). */
enum { FVK_TILES_ALL = 0, FVK_TILES_INTERIOR = 1, FVK_TILES_HALO = 2 };
int fvk_mesh_set_tile_phase(fvk_mesh* mesh, int phase);
/* Diagnostics: out[i] = sum_k in[k][i] (k < nStreams <= 8, in_h = HOST array of device pointers), one element per
 * thread like the operator kernels. tools/stream_probe.py uses it to measure what HBM delivers for N concurrent read
 * streams -- the practical ceiling the 6-8-array operator kernels are compared with next to the 2-stream copy peak. */
int fvk_probe_streams(int nStreams, const double* const* in_h, int64_t n, double* out, fvk_stream stream);
/* tuning switch of the brick kernel (roofline sweeps only): which kernel the override applies to (1 = brick kernel, 3 = affine kernel), threads per block, resident blocks per SM
 * the register allocation aims at; only combinations instantiated in fvk_explicit.cu are honoured (others fall back to
 * the per-cell gather). (0,0,0) restores the defaults. Environment FVK_BRICK_CFG="K,TB,MINB" does the same. */
int fvk_set_brick_config(int kernel, int threads, int minBlocks);
/* A/B switch (roofline harness, parity tests): 0 = the generic brick kernel computes every tile, 1 (default) = tiles of a
 * block-structured mesh whose topology the plan proved affine are computed by the index-free kernel. */
int fvk_set_affine(int enabled);

/* Diagnostics, HOST only (no device needed): build the cell->face stencil and the brick plan of the explicit
 * gather kernel for a mesh description and replay the plan exactly as the kernel reads it. info_h[8] =
 * {nTiles, detected block dims nx ny nz (0 0 0: none -> runs of consecutive cells), brick shape lx by bz,
 * max shared-memory slots per tile}; *badCells_h = number of cells whose replayed (face, sign) sequence differs
 * from the reference's accumulation order (0 = exact). Returns FVK_EUNSUPPORTED when the mesh gets no brick plan
 * (its per-cell face order is not [lower | owned, consecutive ids | boundary]); operators then use the per-cell
 * gather. Environment: FVK_BRICK="lx,by,bz" overrides the default 32,4,4 brick. */
int fvk_brick_plan_selftest(const fvk_mesh_desc* desc_h, int32_t* info_h, int64_t* badCells_h);
/* HOST only: {affine topology proven (0/1), upper side x / y / z is a true boundary (1) or a processor cut (0), number of
 * irregular (boundary / cut layer) cells} of the plan (info_h[5]). */
int fvk_brick_plan_affine_info(const fvk_mesh_desc* desc_h, int32_t* info_h);
/* HOST only: the assumption behind the structured SpMV, checked row by row: result_h[0] = 1 when the topology is affine
 * and every regular row of the SparsityPattern is [c-nx*ny, c-nx, c-1, c, c+1, c+nx, c+nx*ny], 0 when the topology is not
 * affine (generic SpMV is used), -1 on a violation; result_h[1] = rows checked. */
int fvk_brick_plan_structured_rows(const fvk_mesh_desc* desc_h, int64_t* result_h);
/* HOST only: the SparsityPattern exactly as fvk_mesh_create builds it (same host function). result_h[4] = {rows, nnz, every row in
 * stencil order (0/1: what lets the index-free assembly / rAU,HbyA kernels run; holds on every sub-domain of a decomposed block
 * because each half of a row follows the GLOBAL face order), number of offset / diagonal inconsistencies (must be 0)}. */
int fvk_sparsity_selftest(const fvk_mesh_desc* desc_h, int64_t* result_h);

#ifdef __cplusplus
}
#endif
#endif /* FVK_H */
