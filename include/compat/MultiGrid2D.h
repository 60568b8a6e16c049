/* include/compat/MultiGrid2D.h -- shim with the public interface of the reference's MultiGrid2D
   (NOCUDA_TESI/PDE Lyapunov 2D/MultiGrid2D.h:6-37) over libmg_b200.so. */
#ifndef MULTIGRID2D_H
#define MULTIGRID2D_H

#include "Grid2D.h"

class MultiGrid2D
{
	public:
		Grid2D** grids2D;
		int numGrids;

		float* matrixA;
		int sizeA;
		int alfa;
		mg2d_t* engine;
#ifdef MG_COMPAT_CUDA_TESI
		float* d_matrixA; /* CUDA_TESI face (C2/MultiGrid2D.h:12-13): A on the device */
		int sizeX_A;
#endif

		MultiGrid2D(int finestGridSizeXY[], float range[], float* _A, int A_size, int alfa_)
		{
			InitA(_A, A_size, alfa_);
			InitGrids(finestGridSizeXY, range);
		}
		/* the CUDA twin's constructor takes the (square) size as a scalar, CUDA_TESI/CUDA Lyapunov 2D/MultiGrid2D.h:16 */
		MultiGrid2D(int finestGridSize, float range[], float* _A, int A_sizeX, int alfa_)
		{
			int s[2] = {finestGridSize, finestGridSize};
			InitA(_A, A_sizeX, alfa_);
			InitGrids(s, range);
		}
		~MultiGrid2D()
		{
			for (int i = 0; i < numGrids; i++) delete grids2D[i];
			free(grids2D);
			free(matrixA);
#ifdef MG_COMPAT_CUDA_TESI
			cudaFree(d_matrixA);
#endif
			mg2d_destroy(engine);
		}
		void InitA(float* _A, int A_size, int alfa_)
		{
			alfa = alfa_;
			sizeA = A_size;
			matrixA = (float*)malloc((size_t)A_size * A_size * sizeof(float)); // the reference allocates A_size floats and overflows (App. B7)
			for (int i = 0; i < A_size * A_size; i++) matrixA[i] = _A[i];
#ifdef MG_COMPAT_CUDA_TESI
			sizeX_A = A_size;
			MG_CUDA_CHECK(cudaMalloc((void**)&d_matrixA, (size_t)A_size * A_size * sizeof(float)));
			MG_CUDA_CHECK(cudaMemcpy(d_matrixA, matrixA, (size_t)A_size * A_size * sizeof(float), cudaMemcpyHostToDevice));
#endif
		}
		void InitGrids(int finestGridSizeXY[], float range[])
		{
			double r[4], A[4];
			for (int i = 0; i < 4; i++) { r[i] = range[i]; A[i] = matrixA[i]; }
			MG_CHECK(mg2d_create(&engine, finestGridSizeXY, r, A, alfa, MG_F32));
			numGrids = mg2d_num_levels(engine);
			grids2D = (Grid2D**)malloc(numGrids * sizeof(Grid2D*));
			for (int l = 0; l < numGrids; l++) {
				int n = mg2d_level_size(engine, l);
				int s[2] = {n, n};
				grids2D[l] = new Grid2D(s, range, engine, l);
			}
		}

		void Restrict(float* fine, int fsizeXY[], float* coarse, int csizeXY[]) { MG_CHECK(mg2d_restrict_host(engine, fine, fsizeXY, coarse, csizeXY)); }
		void Interpolate(float* fine, int fsizeXY[], float* coarse, int csizeXY[]) { MG_CHECK(mg2d_interpolate_host(engine, fine, fsizeXY, coarse, csizeXY)); }
		void Relax(Grid2D* curGrid, int ncycles)
		{
			int l = level_of(curGrid);
			curGrid->push(engine, l);
			MG_CHECK(mg2d_relax(engine, l, ncycles));
			curGrid->pull(engine, l);
		}
#ifdef MG_COMPAT_CUDA_TESI
		/* CUDA_TESI face (C2/MultiGrid2D.h:20-25): pitched device arrays as (pointer, size, pitch in elements) */
		void Restrict(float* fine, int fsize, int f_pitch, float* coarse, int csize, int c_pitch) { MG_CHECK(mg2d_restrict_device(engine, fine, fsize, f_pitch, coarse, csize, c_pitch)); }
		void Interpolate(float* fine, int fsize, int f_pitch, float* coarse, int csize, int c_pitch) { MG_CHECK(mg2d_interpolate_device(engine, fine, fsize, f_pitch, coarse, csize, c_pitch)); }
		void ApplyCorrection(float* fine, int fineSize, int f_pitch, float* error, int errorSize, int e_pitch) { MG_CHECK(mg2d_apply_correction_device(engine, fine, fineSize, f_pitch, error, errorSize, e_pitch)); }
		void Set(float* v, int size, int pitch, float value, bool modifyBorder) { MG_CHECK(mg2d_set_device(engine, v, size, pitch, value, modifyBorder)); }
		float* CalculateResidual(Grid2D* fine) // caller-owned pitched DEVICE array with the level's pitch (C2/MultiGrid2D.cu:105-127)
		{
			int l = level_of(fine);
			float* d_r = 0;
			MG_CUDA_CHECK(cudaMalloc((void**)&d_r, (size_t)fine->d_pitch * fine->size * sizeof(float)));
			MG_CUDA_CHECK(cudaMemset(d_r, 0, (size_t)fine->d_pitch * fine->size * sizeof(float)));
			MG_CHECK(mg2d_residual_device(engine, l, d_r));
			return d_r;
		}
#else
		float* CalculateResidual(Grid2D* fine)
		{
			int l = level_of(fine);
			fine->push(engine, l);
			float* r = (float*)malloc((size_t)fine->sizeX * fine->sizeY * sizeof(float));
			MG_CHECK(mg2d_residual(engine, l, r));
			return r;
		}
#endif
		void ApplyCorrection(float* fine, int fsizeXY[], float* error, int esizeXY[]) { MG_CHECK(mg2d_apply_correction_host(engine, fine, fsizeXY, error, esizeXY)); }
		void setToValue(float* grid, int sizeXY[], float value, bool modifyBoundaries) { MG_CHECK(mg2d_set_to_value_host(engine, grid, sizeXY, value, modifyBoundaries)); }

		void VCycle(int gridID, int v1, int v2)
		{
			push_all();
			MG_CHECK(mg2d_vcycle(engine, gridID, v1, v2));
			pull_all();
		}
		void FullMultiGridVCycle(int gridID, int v0, int v1, int v2)
		{
			push_all();
			MG_CHECK(mg2d_fmg(engine, gridID, v0, v1, v2));
			pull_all();
		}

		void PrintDiff() { grids2D[0]->PrintDiffApproxReal(mg_compat_open_log("log/diff.txt")); }
		void PrintGrid(int gridID) { grids2D[gridID]->PrintGrid_v(mg_compat_open_log("log/log_v.txt")); }
		void PrintAllGrids_v() { int fd = mg_compat_open_log("log/log_v.txt"); for (int i = 0; i < numGrids; i++) grids2D[i]->PrintGrid_v(fd); }
		void PrintAllGrids_f() { int fd = mg_compat_open_log("log/log_f.txt"); for (int i = 0; i < numGrids; i++) grids2D[i]->PrintGrid_f(fd); }
		void PrintResidual(int) {}
		void PrintMeanAbsoluteError() { grids2D[0]->PrintMeanAbsoluteError(); } // CUDA twin, C2/MultiGrid2D.cu:234-237

	private:
		int level_of(Grid2D* g)
		{
			for (int l = 0; l < numGrids; l++) if (grids2D[l] == g) return l;
			fprintf(stderr, "MultiGrid2D: grid does not belong to this hierarchy\n");
			abort();
		}
		void push_all() { for (int l = 0; l < numGrids; l++) grids2D[l]->push(engine, l); }
		void pull_all() { for (int l = 0; l < numGrids; l++) grids2D[l]->pull(engine, l); }
};
#endif
