// The `cuda` backend as one more sub-project of the reference's sbt build (/root/reference/build.sbt:1-21): copy (or symlink)
// scala/cuda into the reference checkout next to `cpu/` and `gpu/` and add the line below to the root build.sbt — `cuda`
// depends on `Tensors` exactly like `cpu` and `gpu` do (it re-uses Trees, Expressions, NDimensionalAffineTransform, Memory and
// Tensors.TensorBuilder / Tensors.MemoryTrees unchanged; nothing in those projects is modified).
//
//   lazy val cuda = project.dependsOn(Tensors)
//
// and drop this file's settings into cuda/build.sbt. They mirror cpu/build.sbt (same LWJGL 3.2.3 core artifact + natives — the
// dyncall binding CudaNative uses lives in LWJGL's core module — same scalatest) without the lwjgl-opencl dependency.
//
// Not compiled in the environment this repository was developed in (no JVM / sbt / network there): the Scala sources are held
// to the C ABI by tests/test_scala_twin.py, which transliterates CudaTreeWriter's emission order and checks its blobs against
// the C++ mirror's, byte for byte, and against the library's structural cache.

organization := "com.thoughtworks.compute"

name := "cuda"

scalacOptions += "-Ypartial-unification"

libraryDependencies += ("org.lwjgl" % "lwjgl" % "3.2.3").jar().classifier {
  import scala.util.Properties._
  if (isLinux) {
    "natives-linux"
  } else {
    throw new MessageOnlyException(s"libcompute_cuda.so targets Linux + B200 (sm_100a); $osName is not supported")
  }
}

libraryDependencies += "org.lwjgl" % "lwjgl" % "3.2.3"

libraryDependencies += "com.google.guava" % "guava" % "28.2-jre"

libraryDependencies += "com.typesafe.scala-logging" %% "scala-logging" % "3.9.2"

libraryDependencies += "org.scalatest" %% "scalatest" % "3.0.8" % Test

libraryDependencies += "ch.qos.logback" % "logback-classic" % "1.2.3" % Test

addCompilerPlugin("com.github.ghik" %% "silencer-plugin" % "1.4.2")

libraryDependencies += "com.github.ghik" %% "silencer-lib" % "1.4.2"

// libcompute_cuda.so: `python -m compute.scala_b200.build` in this repository writes compute/scala_b200/libcompute_cuda.so
Test / fork := true

Test / javaOptions += s"-Dcom.thoughtworks.compute.cuda.libname=${sys.env.getOrElse("COMPUTE_CUDA_LIB", "libcompute_cuda.so")}"

// golden tree blobs written by the C++ mirror (tests/golden/tree_blobs in this repository), for CudaTreeWriterSpec
Test / javaOptions += s"-Dcom.thoughtworks.compute.cuda.goldens=${sys.env.getOrElse("COMPUTE_CUDA_GOLDENS", "tests/golden/tree_blobs")}"
