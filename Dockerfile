# Image for 8xB200 training with the cgx backend (reference: Dockerfile on nvcr.io/nvidia/pytorch:22.10 + HPC-X MPI).
# No MPI needed here: any launcher (torchrun / srun / mpirun) works.
FROM nvcr.io/nvidia/pytorch:25.03-py3
WORKDIR /opt/torch_cgx_b200
COPY . .
RUN python setup.py build_ext --inplace && pip install --no-build-isolation -e .
ENV CGX_COMPRESSION_QUANTIZATION_BITS=32
CMD ["python", "-m", "pytest", "tests", "-q", "-m", "not gpu"]
